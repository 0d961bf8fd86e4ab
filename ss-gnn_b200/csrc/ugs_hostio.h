// ugs_hostio.h -- shared between ugs_hostio.cu (device-side packing, nvcc) and ugs_hostdeliver.cpp (host side, g++).
#pragma once
#include <stddef.h>
#include <stdint.h>

#define UGS_DELTA_BLOCK 4096  // values per frame-of-reference block of the 16-bit delta format

int ugs_set_error(int code, const char* fmt, ...);

// Byte offsets inside the wire buffer (same on the device and in the pinned host staging buffer) for the PACKED parts of
// the three arrays: pn node ids, pe source columns, pi edge entries.
struct UgsWire {
    size_t nodes;       // pn x (delta16 ? u16 : i32)
    size_t src;         // pe x (delta16 ? u16 : i32)
    size_t ei;          // pairs_u8 ? pi bytes (row0 | row1 << 4) : 2 x pi x i32 (row 0, then row 1)
    size_t nodes_base;  // delta16: one int64 block minimum per 4096 node ids
    size_t src_base;    // delta16: one int64 block minimum per 4096 source columns
    size_t bad;         // two uint64 counters: [0] values outside the 16-bit delta range, [1] values outside u8 pairs / int32
    size_t total;
};

static inline size_t ugs_align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline size_t ugs_delta_blocks(size_t n) { return (n + UGS_DELTA_BLOCK - 1) / UGS_DELTA_BLOCK; }
static inline UgsWire ugs_wire_layout(size_t pn, size_t pe, size_t pi, int pairs_u8, int delta16) {
    UgsWire w;
    const size_t eb = delta16 ? 2 : 4;
    w.nodes = 0;
    w.src = ugs_align256(pn * eb);
    w.ei = w.src + ugs_align256(pe * eb);
    w.nodes_base = w.ei + ugs_align256(pairs_u8 ? pi : pi * 8);
    w.src_base = w.nodes_base + ugs_align256(delta16 ? ugs_delta_blocks(pn) * 8 : 0);
    w.bad = w.src_base + ugs_align256(delta16 ? ugs_delta_blocks(pe) * 8 : 0);
    w.total = w.bad + 256;
    return w;
}

// Layout of the wire buffer when the EMISSION pass writes the edge part itself (ugs_emit_edges_wire): node ids as in the
// packed layout (16-bit deltas per 4096-value block), one byte (row0 | row1 << 4) per edge entry, and the source columns
// as 8-BIT deltas against the base of their 1024-entry block (the first column of the graph that owns the block's first
// entry: 1024 entries are half a ZINC-sized graph's samples, so a block spans two or three graphs' columns).
#define UGS_DELTA8_BLOCK 1024
static inline size_t ugs_delta8_blocks(size_t n) { return (n + UGS_DELTA8_BLOCK - 1) / UGS_DELTA8_BLOCK; }
static inline UgsWire ugs_wire_layout_emit(size_t pn, size_t pe) {
    UgsWire w;
    w.nodes = 0;
    w.src = ugs_align256(pn * 2);
    w.ei = w.src + ugs_align256(pe);
    w.nodes_base = w.ei + ugs_align256(pe);
    w.src_base = w.nodes_base + ugs_align256(ugs_delta_blocks(pn) * 8);
    w.bad = w.src_base + ugs_align256(ugs_delta8_blocks(pe) * 8);
    w.total = w.bad + 256;
    return w;
}

// Device side (ugs_hostio.cu): packs the given parts into d_wire according to *w; zeroes and fills the two counters.
int ugs_pack_parts(const int64_t* nodes, size_t pn, const int64_t* ei_row0, const int64_t* ei_row1, size_t pi,
                   const int64_t* edge_src, size_t pe, int pairs_u8, int delta16, void* d_wire, const UgsWire* w,
                   void* stream);

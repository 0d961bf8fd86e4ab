// Test infrastructure only (oracle/): force-included (-include) when the reference
// samplers are compiled for a host WITHOUT an NVIDIA driver. The reference allocates
// its outputs with `.pinned_memory(true)` (sampler.cpp:190,
// ugs_sampler_batch_extension.cpp:95, epsilon_uniform_sampler.cpp:182), which raises
// "Found no NVIDIA driver" in the CPU-only container. Torch's headers are pulled in
// first (they are `#pragma once`), then the call is rewritten to `pinned_memory(false)`;
// the reference sources themselves are compiled unmodified from where they lie.
#pragma once
#include <torch/extension.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#define pinned_memory(x) pinned_memory(false)

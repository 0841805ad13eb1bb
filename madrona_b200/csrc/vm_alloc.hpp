// vm_alloc.hpp -- growable device allocations with a STABLE base address:
// virtual-address range reserved up front, physical memory mapped behind it in
// granules (cuMemAddressReserve / cuMemCreate / cuMemMap), more mapped later
// without moving anything.  Role of the reference's GPU table storage: 128 GiB
// of VA per column reserved at start-up and committed by a host thread on
// demand (src/mw/device/include/madrona/table.hpp:38-39, src/mw/device/
// state.cpp:29-80, src/mw/cuda_exec.cpp VM allocator thread).  Here growth
// happens between steps (mb2_run checks the tables' high-water marks), so the
// step graph itself never waits for the host.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace mb2 {

struct VMRange {
    void *base = nullptr;
    size_t reserved = 0;
    size_t mapped = 0;
    std::vector<unsigned long long> handles;    // CUmemGenericAllocationHandle
    std::vector<size_t> handleBytes;
};

// reserve `reserve_bytes` of VA on `gpu`, map (and zero) the first `initial_bytes`
bool vmReserve(int gpu, VMRange *r, size_t reserve_bytes, size_t initial_bytes, std::string *err);
// make at least `new_bytes` usable (no-op when already mapped); new memory is zeroed
bool vmGrow(int gpu, VMRange *r, size_t new_bytes, std::string *err);
void vmRelease(VMRange *r);

}

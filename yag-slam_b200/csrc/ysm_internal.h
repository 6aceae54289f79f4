// ysm_internal.h -- shared between the translation units of libysm_b200.so (not part of the C ABI).
#pragma once

// Ends the resident latency kernel (k_match_resident, ysm_resident.cuh) of whichever matcher handle
// owns one on `device`, and waits until it has left the device. Entry points that allocate / free
// device memory, synchronise the device or launch kernels that want the whole machine call this first.
// No-op when no resident kernel is alive.
void ysm_quiesce_device(int device);

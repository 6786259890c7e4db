// host_pack.h -- host-side narrowing of the reference's int32 NCHW tensor to NHWC4 bytes (host_pack.cpp)
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace f8hp {

// rows [r0, r1) of the n*h image rows of x (int32 [n,3,h,w]) -> dst (uint32 per pixel: c0 | c1<<8 | c2<<16).
// Returns false when a value read lies outside [lo, lo + 255] (lo = 0 for an unsigned head, -128 for a signed
// one): the low byte that was stored is then not what the reference's full-int32 head conv would see.
bool pack_rows_nchw_i32(const int32_t *x, uint8_t *dst, int h, int w, long long r0, long long r1, int lo = 0);
// the SIMD body selected on this CPU ("avx512" | "avx2" | "sse2" | "scalar"; F8_HOST_PACK_ISA overrides)
const char *isa_name();
// helper threads per plan: F8_HOST_PACK_THREADS, else min(16, usable cores / LOCAL_WORLD_SIZE); 0 = no host repack
int default_threads();

class Pool {
  public:
    Pool();
    ~Pool();
    Pool(const Pool &) = delete;
    Pool &operator=(const Pool &) = delete;
    // repack rows [r0, r1) with up to `threads` threads (the caller is one of them); returns when done:
    // true when every value lay in [lo, lo + 255]
    bool run(const int32_t *x, uint8_t *dst, int h, int w, long long r0, long long r1, int threads, int lo = 0);

  private:
    struct Impl;
    Impl *p_;
};

}  // namespace f8hp

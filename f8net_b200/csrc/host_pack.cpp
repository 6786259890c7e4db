// host_pack.cpp -- host side of the reference-facing call: narrowing the reference's int32 NCHW tensor
// (/root/reference/fix_train.py:682-692 hands IntModel.forward 8-bit-range integers in int32) to the
// engine-native NHWC4 bytes on the host cores, before the copy to the device.  Plain C++ (compiled by the
// host compiler, no CUDA): AVX-512 / AVX2 / SSE2 bodies selected at run time, a per-plan pool of
// persistent helper threads.  Keeps the low byte of every value, exactly as convert_input_kernel does.
#include "host_pack.h"

#include <immintrin.h>
#include <sched.h>
#include <stdlib.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace f8hp {

namespace {

// Every body also ORs (value - lo) of all it reads into *wide: bits above the low eight mean a value
// outside [lo, lo + 255], which the reference's head conv (full int32) would not treat as its low byte.
void pack_row_scalar(const int32_t *c0, const int32_t *c1, const int32_t *c2, uint32_t *o, int i, int w, uint32_t lo,
                     uint32_t *wide) {
    uint32_t acc = 0;
    for (; i < w; ++i) {
        const uint32_t a = (uint32_t)c0[i], b = (uint32_t)c1[i], c = (uint32_t)c2[i];
        acc |= (a - lo) | (b - lo) | (c - lo);
        o[i] = (a & 0xffu) | ((b & 0xffu) << 8) | ((c & 0xffu) << 16);
    }
    *wide |= acc;
}

// streaming stores: the staging is written once and read by the DMA engine, never by this core
__attribute__((target("sse2"))) int pack_row_sse2(const int32_t *c0, const int32_t *c1, const int32_t *c2, uint32_t *o, int w,
                                                  uint32_t lo, uint32_t *wide) {
    int i = 0;
    if ((reinterpret_cast<uintptr_t>(o) & 15u) != 0) return 0;
    const __m128i m = _mm_set1_epi32(0xff), vlo = _mm_set1_epi32((int)lo);
    __m128i acc = _mm_setzero_si128();
    for (; i + 4 <= w; i += 4) {
        const __m128i ra = _mm_loadu_si128(reinterpret_cast<const __m128i *>(c0 + i));
        const __m128i rb = _mm_loadu_si128(reinterpret_cast<const __m128i *>(c1 + i));
        const __m128i rc = _mm_loadu_si128(reinterpret_cast<const __m128i *>(c2 + i));
        acc = _mm_or_si128(acc, _mm_or_si128(_mm_sub_epi32(ra, vlo), _mm_or_si128(_mm_sub_epi32(rb, vlo), _mm_sub_epi32(rc, vlo))));
        const __m128i a = _mm_and_si128(ra, m), b = _mm_and_si128(rb, m), c = _mm_and_si128(rc, m);
        _mm_stream_si128(reinterpret_cast<__m128i *>(o + i), _mm_or_si128(a, _mm_or_si128(_mm_slli_epi32(b, 8), _mm_slli_epi32(c, 16))));
    }
    alignas(16) uint32_t lanes[4];
    _mm_store_si128(reinterpret_cast<__m128i *>(lanes), acc);
    *wide |= lanes[0] | lanes[1] | lanes[2] | lanes[3];
    return i;
}

__attribute__((target("avx2"))) int pack_row_avx2(const int32_t *c0, const int32_t *c1, const int32_t *c2, uint32_t *o, int w,
                                                  uint32_t lo, uint32_t *wide) {
    int i = 0;
    if ((reinterpret_cast<uintptr_t>(o) & 31u) != 0) return 0;
    const __m256i m = _mm256_set1_epi32(0xff), vlo = _mm256_set1_epi32((int)lo);
    __m256i acc = _mm256_setzero_si256();
    for (; i + 8 <= w; i += 8) {
        const __m256i ra = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(c0 + i));
        const __m256i rb = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(c1 + i));
        const __m256i rc = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(c2 + i));
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_sub_epi32(ra, vlo),
                                                   _mm256_or_si256(_mm256_sub_epi32(rb, vlo), _mm256_sub_epi32(rc, vlo))));
        const __m256i a = _mm256_and_si256(ra, m), b = _mm256_and_si256(rb, m), c = _mm256_and_si256(rc, m);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(o + i),
                            _mm256_or_si256(a, _mm256_or_si256(_mm256_slli_epi32(b, 8), _mm256_slli_epi32(c, 16))));
    }
    alignas(32) uint32_t lanes[8];
    _mm256_store_si256(reinterpret_cast<__m256i *>(lanes), acc);
    for (int k = 0; k < 8; ++k) *wide |= lanes[k];
    return i;
}

__attribute__((target("avx512f"))) int pack_row_avx512(const int32_t *c0, const int32_t *c1, const int32_t *c2, uint32_t *o, int w,
                                                       uint32_t lo, uint32_t *wide) {
    int i = 0;
    if ((reinterpret_cast<uintptr_t>(o) & 63u) != 0) return 0;
    const __m512i m = _mm512_set1_epi32(0xff), vlo = _mm512_set1_epi32((int)lo);
    __m512i acc = _mm512_setzero_si512();
    for (; i + 16 <= w; i += 16) {
        const __m512i ra = _mm512_loadu_si512(c0 + i), rb = _mm512_loadu_si512(c1 + i), rc = _mm512_loadu_si512(c2 + i);
        acc = _mm512_or_si512(acc, _mm512_or_si512(_mm512_sub_epi32(ra, vlo),
                                                   _mm512_or_si512(_mm512_sub_epi32(rb, vlo), _mm512_sub_epi32(rc, vlo))));
        const __m512i a = _mm512_and_si512(ra, m), b = _mm512_and_si512(rb, m), c = _mm512_and_si512(rc, m);
        _mm512_stream_si512(reinterpret_cast<__m512i *>(o + i),
                            _mm512_or_si512(a, _mm512_or_si512(_mm512_slli_epi32(b, 8), _mm512_slli_epi32(c, 16))));
    }
    alignas(64) uint32_t lanes[16];
    _mm512_store_si512(lanes, acc);
    for (int k = 0; k < 16; ++k) *wide |= lanes[k];
    return i;
}

int isa_level() {
    static const int level = [] {
        if (const char *e = getenv("F8_HOST_PACK_ISA")) return atoi(e);      // 0 scalar, 1 SSE2, 2 AVX2, 3 AVX-512
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx512f")) return 3;
        if (__builtin_cpu_supports("avx2")) return 2;
        if (__builtin_cpu_supports("sse2")) return 1;
        return 0;
    }();
    return level;
}

}  // namespace

const char *isa_name() {
    static const char *names[] = {"scalar", "sse2", "avx2", "avx512"};
    const int l = isa_level();
    return names[l < 0 ? 0 : (l > 3 ? 3 : l)];
}

bool pack_rows_nchw_i32(const int32_t *x, uint8_t *dst, int h, int w, long long r0, long long r1, int lo_value) {
    const size_t hw = (size_t)h * w;
    const int level = isa_level();
    const uint32_t lo = (uint32_t)lo_value;
    uint32_t wide = 0;
    for (long long r = r0; r < r1; ++r) {
        const long long img = r / h;
        const int y = (int)(r - img * h);
        const int32_t *c0 = x + (size_t)img * 3 * hw + (size_t)y * w;
        const int32_t *c1 = c0 + hw, *c2 = c1 + hw;
        uint32_t *o = reinterpret_cast<uint32_t *>(dst) + (size_t)r * w;
        int i = 0;
        if (level >= 3) i = pack_row_avx512(c0, c1, c2, o, w, lo, &wide);
        if (level >= 2 && i < w - 7) i += pack_row_avx2(c0 + i, c1 + i, c2 + i, o + i, w - i, lo, &wide);
        if (level >= 1 && i < w - 3) i += pack_row_sse2(c0 + i, c1 + i, c2 + i, o + i, w - i, lo, &wide);
        pack_row_scalar(c0, c1, c2, o, i, w, lo, &wide);
    }
    if (level >= 1) _mm_sfence();
    return (wide & ~0xffu) == 0;
}

int default_threads() {
    static const int t = [] {
        if (const char *e = getenv("F8_HOST_PACK_THREADS")) return std::max(0, atoi(e));   // 0 = ship the int32 tensor as is
        // the cores this process may run on, shared by the ranks of the box (torchrun sets LOCAL_WORLD_SIZE):
        // eight ranks with sixteen helpers each on 32 cores only fight for the same memory controllers
        int cores = 0;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        if (cores <= 0) cores = (int)std::thread::hardware_concurrency();
        int ranks = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        return std::min(16, std::max(1, cores / ranks));
    }();
    return t;
}

// ---------------------------------------------------------------------------------------------
// Persistent helper threads (spawning threads per call costs a fifth of the repack itself).  Workers
// sleep on a condition variable between calls; the pool is owned by its plan (no process-wide lock:
// two plans on two streams repack concurrently) and joined when the plan is destroyed.
// ---------------------------------------------------------------------------------------------
struct Pool::Impl {
    std::mutex call, m;
    std::condition_variable work, done;
    std::vector<std::thread> threads;
    const int32_t *x = nullptr;
    uint8_t *dst = nullptr;
    int h = 0, w = 0, T = 0, pending = 0, lo = 0;
    bool in_range = true;
    long long r0 = 0, rows = 0;
    unsigned long long gen = 0;
    bool quit = false;

    void worker(int id) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m);
            work.wait(lk, [&] { return gen != seen || quit; });
            if (quit) return;
            seen = gen;
            if (id >= T) continue;                         // not needed for this call
            const int32_t *xx = x;
            uint8_t *dd = dst;
            const int hh = h, ww = w, TT = T, llo = lo;
            const long long rr0 = r0, n = rows;
            lk.unlock();
            const bool ok = pack_rows_nchw_i32(xx, dd, hh, ww, rr0 + n * id / TT, rr0 + n * (id + 1) / TT, llo);
            lk.lock();
            if (!ok) in_range = false;
            if (--pending == 0) done.notify_one();
        }
    }
};

Pool::Pool() : p_(new Impl) {}

Pool::~Pool() {
    {
        std::lock_guard<std::mutex> lk(p_->m);
        p_->quit = true;
    }
    p_->work.notify_all();
    for (auto &t : p_->threads) t.join();
    delete p_;
}

bool Pool::run(const int32_t *x, uint8_t *dst, int h, int w, long long r0, long long r1, int threads, int lo) {
    const long long rows = r1 - r0;
    if (rows <= 0) return true;
    const int T = (int)std::min<long long>(std::max(1, threads), std::max<long long>(1, rows / 64));
    std::lock_guard<std::mutex> serial(p_->call);          // one repack at a time per pool
    if (T > 1) {
        std::unique_lock<std::mutex> lk(p_->m);
        while ((int)p_->threads.size() < T - 1) {
            const int id = (int)p_->threads.size() + 1;
            p_->threads.emplace_back(&Impl::worker, p_, id);
        }
        p_->x = x; p_->dst = dst; p_->h = h; p_->w = w; p_->r0 = r0; p_->rows = rows; p_->T = T; p_->lo = lo;
        p_->in_range = true;
        p_->pending = T - 1;
        ++p_->gen;
        lk.unlock();
        p_->work.notify_all();
    }
    bool ok = pack_rows_nchw_i32(x, dst, h, w, r0, r0 + rows / T, lo);
    if (T > 1) {
        std::unique_lock<std::mutex> lk(p_->m);
        p_->done.wait(lk, [&] { return p_->pending == 0; });
        ok = ok && p_->in_range;
    }
    return ok;
}

}  // namespace f8hp

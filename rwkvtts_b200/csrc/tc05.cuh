// Blackwell (sm_100a) primitives used by the WKV-7 kernels: tcgen05 MMA (tf32), tensor memory,
// mbarriers, proxy fences.  Inline PTX only; nothing here depends on CUTLASS.
//
// Shared-memory operand layout (no swizzle, "interleaved" canonical form): the unit is a core
// matrix of 8 rows x 16 bytes (4 tf32), stored as 128 contiguous bytes.
//   K-major operand  X[r][k]  (r = M or N index, k contiguous in a row of the core matrix):
//       byte(r,k) = (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4
//   MN-major operand X[k][r]  (r = M or N index contiguous inside the core matrix row):
//       byte(r,k) = (r/4)*SBO + (k/8)*LBO + (k%8)*16 + (r%4)*4
// One tf32 MMA consumes K = 8: two core matrices along K for a K-major operand, one 8-row group
// for an MN-major operand.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rwkvtts {
namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- descriptors ---------------------------------------------------------------------------
// instruction descriptor, kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                 // D format  = F32
         | (2u << 7)                 // A format  = TF32
         | (2u << 10)                // B format  = TF32
         | ((a_mn_major ? 1u : 0u) << 15)
         | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor, no swizzle; byte offsets must be multiples of 16
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4)
         | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16)
         | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32)
         | (1ull << 46);             // descriptor version 1 (sm_100)
}

// ---- MMA issue ---------------------------------------------------------------------------------
// Call these from a fully converged warp with warp-uniform arguments, inside `if (elect_one())`:
// a thread-divergent guard (if (tid == 0)) makes ptxas wrap every UTCHMMA in an election loop
// (~45 cycles per instruction, measured).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

// ---- tensor memory --------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 32 bit, N consecutive columns: thread i of warp w reads lane 32*(w%4)+i
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
           "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
           "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
           "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
           "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
           "r"(__float_as_uint(v[15])) : "memory");
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// one arrival per warp: every arrival wakes all warps suspended on the barrier (they re-check the phase and go
// back to sleep, 4 instructions each time), so per-thread arrivals on a 256-count barrier cost the waiters ~75
// wake-ups per hand-off (measured with ncu: 17-40 % of all issued instructions).  Writers fence (proxy / tcgen05)
// themselves before calling this; __syncwarp orders their writes before lane 0's release-arrive.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t *bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) expires;
// without a hint the default limit is short and 20+ waiting warps burn a third of the issue slots
// re-polling (measured with ncu on the backward kernel).
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    return ok != 0;
}

// ---- watchdog ------------------------------------------------------------------------------------
// Every mbarrier wait of the chunked kernels is bounded: a hand-off that has not arrived after
// 2.2 - 4.4 s (two ticks of the upper word of the SM cycle counter; a whole launch takes < 2 ms) writes one record
//   {kernel id, barrier offset in dynamic shared memory, parity, block, thread, time waited}   (one per stuck warp)
// to a pinned host buffer (capi.cu owns it; rwkvtts_watchdog_report() formats it) and traps, so a
// protocol error surfaces as a CUDA error with a diagnosis instead of a device that spins forever.
constexpr unsigned long long kWatchdogMagic = 0x57444f4752574b56ull;   // "WDOGRWKV"
static __device__ unsigned long long *g_wd_rec = nullptr;              // one copy per translation unit
static __device__ unsigned int g_wd_kernel = 0;
// record layout (u64 words): [0] magic once any entry exists, entries from word 8: ONE packed word per distinct barrier
// (barriers are consecutive 8-byte words of the kernel's shared-memory struct, so offset / 8 spreads them over the
// entries; thousands of warps are stuck at once and a deadlock is a cycle of waits -- the record keeps one waiter of
// every barrier somebody is stuck on):  valid(1) | kernel(3) | parity(1) | warp(6) | block(24) | smem offset(24)
constexpr int kWatchdogEntries = 32, kWatchdogWords = 8 + kWatchdogEntries;
// The report lives in ONE out-of-line function per translation unit that never returns: a call that does not come back
// needs no register saved around it (an ordinary out-of-line report function cost the forward 40 % in local-memory
// traffic at its ~40 call sites, and the fully inlined report 21-24 % more code -- both measured), so a wait site is
// two polls, one read of the upper clock word and a compare.
[[noreturn]] static __device__ __noinline__ void mbar_die(uint32_t bar_addr, uint32_t parity) {
    extern __shared__ __align__(128) unsigned char wd_dyn_smem[];
    unsigned long long *r = g_wd_rec;
    if (r != nullptr) {
        const uint32_t off = bar_addr - smem_u32(wd_dyn_smem);
        r[8 + ((off >> 3) % kWatchdogEntries)] = (1ull << 63) | ((unsigned long long)g_wd_kernel << 60) |
                                                 ((unsigned long long)(parity & 1u) << 59) |
                                                 ((unsigned long long)(threadIdx.x >> 5) << 48) |
                                                 ((unsigned long long)(blockIdx.x & 0xffffffu) << 24) | (off & 0xffffffu);
        r[0] = kWatchdogMagic;
        __threadfence_system();
    }
    // nobody traps at once: a trap aborts the grid, stores in flight included, and the other stuck warps run out of
    // patience within a few ms of this one
    const long long t1 = clock64();
    while (clock64() - t1 < 100000000ll) {}
    __trap();
    for (;;) {}
}
__device__ __forceinline__ uint32_t clock_hi() {        // upper word of the SM cycle counter: one tick = 2^32 cycles ~ 2.2 s
    uint32_t t;
    asm volatile("mov.u32 %0, %%clock_hi;" : "=r"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    // hot path = the unbounded loop's first two polls: try_wait suspends the thread in hardware until the phase
    // completes or its 10 ms hint expires, so a third poll only happens when something is already badly late.  Bound:
    // two ticks of the upper clock word, i.e. 2.2 - 4.4 s (reading %globaltimer instead costs ~1 us: taken on every
    // blocking wait it slowed the forward kernel by 17 %, measured).
    if (mbar_try_wait(bar, parity)) return;
    if (mbar_try_wait(bar, parity)) return;
    const uint32_t t0 = clock_hi();
    while (!mbar_try_wait(bar, parity)) {
        if (clock_hi() - t0 >= 2u) mbar_die(smem_u32(bar), parity);
    }
}
// host side: point this translation unit's record pointer at the pinned buffer (once per device)
inline cudaError_t watchdog_install(unsigned long long *rec, unsigned int kernel_id) {
    cudaError_t e = cudaMemcpyToSymbol(g_wd_rec, &rec, sizeof(rec));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(g_wd_kernel, &kernel_id, sizeof(kernel_id));
}

// ---- bulk asynchronous copies (TMA engine, 1-D): no registers, one instruction per contiguous piece ------------
// shared -> global, tracked by the issuing thread's bulk-group; sizes / addresses multiples of 16 bytes.
// Writers of the shared source execute fence_proxy_async() and synchronise with the issuer first.
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// global -> shared, completion counted in bytes on an mbarrier (pair with mbar_expect_tx)
__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---- tensor-map copies (TMA engine, tiled): one box of a [B*T, H, 64] bf16 activation tensor per instruction ----
// `tmap` is the address of a CUtensorMap that lives in a __grid_constant__ kernel parameter (tma_map.h builds it);
// the box lands densely ([tokens][64] bf16, 128-byte rows, no swizzle) at `sdst` (128-byte aligned) and its bytes are
// counted on `bar` (pair with mbar_expect_tx).  Coordinates, innermost first: channel, head, token.  Tokens beyond the
// tensor's end are filled with zeros (and still counted).
__device__ __forceinline__ void tma_load_box3(void *sdst, const void *tmap, int c_chan, int c_head, int c_tok, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        :: "r"(smem_u32(sdst)), "l"(tmap), "r"(c_chan), "r"(c_head), "r"(c_tok), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(tmap) : "memory");
}

// ---- cp.async (LDGSTS): global -> shared without staging registers ------------------------------
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(smem)), "l"(gmem) : "memory");
}
// the same copy with a source size: 0 bytes are read and the 8 destination bytes are zero-filled when !valid (tokens
// beyond the end of a packed sequence)
__device__ __forceinline__ void cp_async8_zfill(void *smem, const void *gmem, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(smem_u32(smem)), "l"(gmem), "r"(valid ? 8u : 0u) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// ex2.approx.ftz: exp2 without the denormal fix-up code that __expf drags in
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// element offsets (in floats) of the canonical no-swizzle layouts described at the top
__host__ __device__ constexpr int kmajor_off(int r, int k, int lbo_f, int sbo_f) {
    return (r >> 3) * sbo_f + (k >> 2) * lbo_f + (r & 7) * 4 + (k & 3);
}
__host__ __device__ constexpr int mnmajor_off(int r, int k, int lbo_f, int sbo_f) {
    return (r >> 2) * sbo_f + (k >> 3) * lbo_f + (k & 7) * 4 + (r & 3);
}

}  // namespace tc05
}  // namespace rwkvtts

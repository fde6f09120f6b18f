// tools/umma_bench.cu -- microbenchmark + semantics probe of tcgen05.mma operand layouts on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench tools/umma_bench.cu && tools/umma_bench
// (1) cycles per MMA (M = 128, K = 16, bf16) for N in {64, 128, 192, 256} with the A/B operands in shared memory in
//     the canonical K-major layouts: no swizzle ("interleave"), 32 B, 64 B and 128 B swizzle;
// (2) correctness of a 128-byte-swizzled A operand whose descriptor start is shifted by whole rows (start address
//     + 128 * s, base_offset = (start >> 7) & 7) -- the trick the residual-tower kernel needs for its x taps.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t start, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((start >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7u) << 49;
    d |= (uint64_t)(layout & 7u) << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

// layout: 0 none, 6 = 32B, 4 = 64B, 2 = 128B.  out[0] = cycles for `iters` MMAs.
__global__ void __launch_bounds__(384, 1) k_bench(int layout, int n, int iters, int a_shift_bytes, long long* out, int pattern = 0, int pollers = 0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, never;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { stop = 0; mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&never), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        // A: 136 rows of K = 64 (so a shifted start stays inside), B: 256 rows of K = 64
        uint32_t row_bytes = layout == 2 ? 128 : layout == 4 ? 64 : layout == 6 ? 32 : 16;
        uint32_t lbo, sbo;
        const uint32_t a0 = base + 1024, b0 = base + 64 * 1024;
        if (layout == 0) { sbo = 128; lbo = 4096; }           // K halves 4 KB apart, rows 16 bytes apart
        else { sbo = 8 * row_bytes; lbo = 16; }               // swizzled K-major: LBO unused for one atom in K
        const uint32_t idesc = make_idesc((uint32_t)n);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t k = (uint32_t)(i & 3);             // walk the four K steps of a 64-wide row like a real kernel
            uint32_t as = a0 + (uint32_t)a_shift_bytes, bs = b0;
            if (layout == 0) { as += 2 * k * lbo; bs += 2 * k * lbo; }
            else if (layout == 2) { as += 32 * k; bs += 32 * k; }
            else if (layout == 4) { as += 32 * (k & 1); bs += 32 * (k & 1); }
            const uint64_t ad = make_desc(as, lbo, sbo, (uint32_t)layout, layout == 2 ? ((as >> 7) & 7u) : 0u);
            const uint64_t bd = make_desc(bs, lbo, sbo, (uint32_t)layout, 0);
            uint32_t dcol;
            if (pattern == 0) dcol = (uint32_t)((i & 1) * 256);
            else if (pattern == 1) dcol = 64u * (uint32_t)(i % 4);                 // sliding, partially overlapping ranges
            else if (pattern == 2) dcol = 0;                                      // always the same accumulator
            else if (pattern == 3) { const int hp[6] = {0, 3, 1, 4, 2, 5}; const int h = hp[i % 6]; dcol = 64u * (uint32_t)(h > 0 ? h - 1 : 0); }
            else if (pattern == 4) dcol = 64u * (uint32_t)(i % 6);                // N = 64: six distinct blocks in turn
            else dcol = 64u * (uint32_t)((i >> 2) % 4);                           // four MMAs on one range, then slide by 64
            umma(tmem + dcol, ad, bd, idesc, 1u);
        }
        umma_commit(smem_u32(&bar));
        while (!mbar_try_wait(smem_u32(&bar), 0)) {}
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
        stop = 1;
    } else if (threadIdx.x >= 128 && (int)threadIdx.x < 128 + pollers) {
        // idle warps parked on an mbarrier that never completes, like epilogue warps waiting for an accumulator
        while (!stop) { if (mbar_try_wait(smem_u32(&never), 0)) break; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// Semantics: A rows r (0..143) of K = 64 bf16, 128-byte swizzled on ABSOLUTE shared addresses; A[r][k] = mode ? k : r.
// B = 64 x 64 identity (128-byte swizzled).  D = A(shifted by `shift` rows) x B^T, M = 128, N = 64, four K steps.
// out[r * 64 + n] = D[r][n].
__global__ void __launch_bounds__(128, 1) k_shift(int shift, int mode, int use_base_off, float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    unsigned char* g = smem + (base - smem_u32(smem));
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0;
    __syncthreads();
    const uint32_t a_off = 2048, b_off = 32 * 1024;  // A row 0 at base + 2048 (1024-aligned), rows -8.. are zero
    for (int i = threadIdx.x; i < 144 * 64; i += blockDim.x) {
        const int r = i / 64, k = i % 64;
        const uint32_t addr = a_off + (uint32_t)r * 128u;                       // absolute row address (relative to the 1024-aligned base)
        const uint32_t chunk = ((uint32_t)k >> 3) ^ ((addr >> 7) & 7u);
        *reinterpret_cast<__nv_bfloat16*>(g + addr + chunk * 16 + (k & 7) * 2) = __float2bfloat16(mode ? (float)k : (float)r);
    }
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int n = i / 64, k = i % 64;
        const uint32_t addr = b_off + (uint32_t)n * 128u;
        const uint32_t chunk = ((uint32_t)k >> 3) ^ ((addr >> 7) & 7u);
        *reinterpret_cast<__nv_bfloat16*>(g + addr + chunk * 16 + (k & 7) * 2) = __float2bfloat16(n == k ? 1.0f : 0.0f);
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) {
            const uint32_t as = base + a_off + (uint32_t)(shift * 128) + 32u * k, bs = base + b_off + 32u * k;
            const uint64_t ad = make_desc(as, 16, 1024, 2, use_base_off ? ((as >> 7) & 7u) : 0u);
            const uint64_t bd = make_desc(bs, 16, 1024, 2, 0);
            umma(tmem, ad, bd, make_idesc(64), k > 0);
        }
        umma_commit(smem_u32(&bar));
        while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int c = 0; c < 64; c += 8) {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int e = 0; e < 8; ++e) out[(32 * warp + lane) * 64 + c + e] = __uint_as_float(r[e]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
    long long* d_out;
    CK(cudaMalloc(&d_out, 1024 * sizeof(long long)));
    const int smem = 161 * 1024 + 1024;
    CK(cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 4096;
    const int layouts[4] = {0, 6, 4, 2};
    const char* names[4] = {"none(interleave)", "swizzle 32B", "swizzle 64B", "swizzle 128B"};
    for (int grid : {1}) {
        for (int li = 0; li < 4; ++li)
            for (int n : {64, 128, 192, 256}) {
                for (int shift : {0, 16}) {
                    if (shift && layouts[li] != 0 && layouts[li] != 2) continue;
                    const int sb = layouts[li] == 2 && shift ? 128 : shift;  // one row: 16 bytes unswizzled, 128 bytes in 128B swizzle
                    k_bench<<<grid, 384, smem>>>(layouts[li], n, iters, sb, d_out, 0, 0);
                    CK(cudaDeviceSynchronize());
                    std::vector<long long> h(grid);
                    CK(cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                    long long mx = 0;
                    for (auto v : h) mx = v > mx ? v : mx;
                    printf("grid %3d  %-18s N=%3d a_shift=%3dB : %.1f cycles/MMA (floor %d)\n", grid, names[li], n, sb, (double)mx / iters, n / 2);
                }
            }
    }
    {
        const char* pn[6] = {"two accumulators alternating", "ranges sliding by 64 columns every MMA", "one accumulator", "h' order 0,3,1,4,2,5",
                             "six distinct 64-column blocks in turn", "four MMAs per range, then slide by 64"};
        for (int pollers : {0, 256})
            for (int pat = 0; pat < 6; ++pat)
                for (int n : {64, 192}) {
                    if ((pat == 4) != (n == 64) && pat == 4) continue;
                    if (n == 64 && pat != 4 && pat != 0) continue;
                    k_bench<<<1, 384, smem>>>(0, n, iters, 0, d_out, pat, pollers);
                    CK(cudaDeviceSynchronize());
                    long long h1;
                    CK(cudaMemcpy(&h1, d_out, sizeof h1, cudaMemcpyDeviceToHost));
                    printf("no swizzle N=%3d pollers=%3d pattern %d (%s): %.1f cycles/MMA\n", n, pollers, pat, pn[pat], (double)h1 / iters);
                }
    }
    // ---- semantics of shifted 128B-swizzled starts
    float* d_f;
    CK(cudaMalloc(&d_f, 128 * 64 * sizeof(float)));
    CK(cudaFuncSetAttribute(k_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024));
    std::vector<float> hf(128 * 64);
    for (int use_bo : {1, 0})
        for (int shift : {0, 1, -1, 3, 8, 5}) {
            int bad_rows = 0, bad_cols = 0;
            for (int mode = 0; mode < 2; ++mode) {
                k_shift<<<1, 128, 66 * 1024>>>(shift, mode, use_bo, d_f);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(hf.data(), d_f, hf.size() * sizeof(float), cudaMemcpyDeviceToHost));
                for (int r = 0; r < 128; ++r)
                    for (int n = 0; n < 64; ++n) {
                        const int src = r + shift;
                        const float want = (src < 0) ? 0.0f : (mode ? (float)n : (float)src);
                        if (hf[r * 64 + n] != want) { if (mode) ++bad_cols; else ++bad_rows; }
                    }
                if (mode == 0 && bad_rows) printf("   e.g. D[1][0]=%g D[8][0]=%g D[9][0]=%g D[17][3]=%g\n", hf[64], hf[8 * 64], hf[9 * 64], hf[17 * 64 + 3]);
            }
            printf("shift %+d rows, base_offset %s: row test mismatches %d, column test mismatches %d\n", shift, use_bo ? "set" : "0", bad_rows, bad_cols);
        }
    printf("done\n");
    return 0;
}

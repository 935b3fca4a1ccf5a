// umma_probe.cu -- development probe (not product code): one CTA, one tcgen05.mma chain D[128 x N] = A[128 x K] * B[N x K]^T
// with both operands in shared memory under caller-supplied descriptors, to pin the MN-major (transposed) no-swizzle layout
// the weight-gradient kernel relies on.   nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libumma_probe.so
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

extern "C" __global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const unsigned char* a_img, int a_bytes, const unsigned char* b_img, int b_bytes, int N, int ksteps,
                  uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kadv, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kadv,
                  uint32_t idesc, float* out /* 128 x N */, int reps, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    unsigned char* sa = smem;
    unsigned char* sb = smem + ((a_bytes + 1023) / 1024) * 1024;
    for (int i = threadIdx.x; i < a_bytes / 16; i += 128) ((uint4*)sa)[i] = ((const uint4*)a_img)[i];
    for (int i = threadIdx.x; i < b_bytes / 16; i += 128) ((uint4*)sb)[i] = ((const uint4*)b_img)[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const long long c0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
        for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t da = make_desc(smem_u32(sa) + kk * a_kadv, a_lbo, a_sbo);
            const uint64_t db = make_desc(smem_u32(sb) + kk * b_kadv, b_lbo, b_sbo);
            const uint32_t acc = (kk > 0) || (rep > 0 && rep + 1 < reps);  // last repetition restarts the accumulation: out = one clean product
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile(
            "{\n\t.reg .pred p;\n\tW2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE2;\n\tbra W2;\n\tDONE2:\n\t}\n" ::"r"(smem_u32(&bar))
            : "memory");
        if (cycles) *cycles = clock64() - c0;
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE;\n\tbra W;\n\tDONE:\n\t}\n" ::"r"(smem_u32(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(lane_addr + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[(size_t)threadIdx.x * N + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

extern "C" int umma_probe(const unsigned char* a_img, int a_bytes, const unsigned char* b_img, int b_bytes, int N, int ksteps,
                          uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kadv, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kadv,
                          uint32_t idesc, float* out, int reps, long long* cycles) {
    const int smem = ((a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
    cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_probe_kernel<<<1, 128, smem>>>(a_img, a_bytes, b_img, b_bytes, N, ksteps, a_lbo, a_sbo, a_kadv, b_lbo, b_sbo, b_kadv, idesc, out, reps, cycles);
    cudaError_t e = cudaDeviceSynchronize();
    return (int)e;
}

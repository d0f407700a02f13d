// Stand-alone bring-up test of the tcgen05 int8 path used by the tensor-core local kernel (csrc/local_tc.cu):
// C[128 x 256] (s32, TMEM) = A[128 x K] (s8, smem, K-major) * B[256 x K]^T (s8, smem, K-major), operands written into
// shared memory by the threads themselves in the canonical no-swizzle K-major layout (8 x 16-byte core matrices), one
// thread issuing tcgen05.mma.kind::i8, completion through tcgen05.commit -> mbarrier, accumulator read back with
// tcgen05.ld.32x32b.  Every wait is bounded: a wrong descriptor ends in FAIL, not in a hang.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_i8_test umma_i8_test.cu && ./umma_i8_test
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int TM = 128, TN = 256, KC = 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48),
    // layout type SWIZZLE_NONE = 0 [61,64)
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// InstrDescriptor: c_format S32 = 2 [4,6), a/b format signed 8 bit = 1 [7,10) [10,13), K-major both, N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); spin++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// canonical K-major no-swizzle tile with KC bytes of K per row: element (r, k)
__device__ __forceinline__ uint32_t canon(uint32_t r, uint32_t k) { return (r >> 3) * (KC / 16 * 128) + (k >> 4) * 128 + (r & 7) * 16 + (k & 15); }

__global__ void __launch_bounds__(128) k_umma_test(const int8_t *A, const int8_t *B, int K, int32_t *C, int *err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem, *sB = smem + TM * KC;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0;
    bool ok = true;
    for (int k0 = 0; k0 < K; k0 += KC) {
        for (int i = tid; i < TM * KC; i += 128) sA[canon(i / KC, i % KC)] = (uint8_t)A[(size_t)(i / KC) * K + k0 + i % KC];
        for (int i = tid; i < TN * KC; i += 128) sB[canon(i / KC, i % KC)] = (uint8_t)B[(size_t)(i / KC) * K + k0 + i % KC];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int kk = 0; kk < KC / 32; kk++) {
                const uint64_t da = umma_desc(smem_u32(sA) + kk * 256, 128, KC / 16 * 128);
                const uint64_t db = umma_desc(smem_u32(sB) + kk * 256, 128, KC / 16 * 128);
                const uint32_t acc = (k0 > 0 || kk > 0) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(IDESC), "r"(acc), "r"(0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
        }
        ok = mbar_wait(smem_u32(&mbar), phase) && ok;
        phase ^= 1;
        if (!__syncthreads_and(ok)) break;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (ok) {
        for (int c = 0; c < TN; c += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                         "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                           "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                           "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; j++) C[(size_t)(warp * 32 + lane) * TN + c + j] = (int32_t)v[j];
        }
    } else if (tid == 0) {
        *err = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

int main() {
    const int K = 256;
    std::vector<int8_t> A(TM * K), B(TN * K);
    srand(1);
    for (auto &x : A) x = (rand() % 3) == 0;
    for (auto &x : B) x = (rand() % 4) == 0;
    std::vector<int32_t> ref(TM * TN, 0), got(TM * TN, -1);
    for (int m = 0; m < TM; m++)
        for (int n = 0; n < TN; n++) {
            int s = 0;
            for (int k = 0; k < K; k++) s += A[m * K + k] * B[n * K + k];
            ref[m * TN + n] = s;
        }
    int8_t *dA, *dB;
    int32_t *dC;
    int *dErr, hErr = 0;
    cudaMalloc(&dA, A.size());
    cudaMalloc(&dB, B.size());
    cudaMalloc(&dC, got.size() * 4);
    cudaMalloc(&dErr, 4);
    cudaMemset(dErr, 0, 4);
    cudaMemset(dC, 0xff, got.size() * 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    const size_t smem = (TM + TN) * KC + 1024;
    cudaFuncSetAttribute(k_umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_umma_test<<<1, 128, smem>>>(dA, dB, K, dC, dErr);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(got.data(), dC, got.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&hErr, dErr, 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (size_t i = 0; i < ref.size(); i++) bad += ref[i] != got[i];
    printf("cuda: %s, barrier timeout: %d, mismatches: %ld of %zu (C[0][0..3] = %d %d %d %d, ref %d %d %d %d; C[5][7] = %d ref %d)\n",
           cudaGetErrorString(e), hErr, bad, ref.size(), got[0], got[1], got[2], got[3], ref[0], ref[1], ref[2], ref[3], got[5 * TN + 7], ref[5 * TN + 7]);
    printf("%s\n", (e == cudaSuccess && !hErr && bad == 0) ? "PASS" : "FAIL");
    return (e == cudaSuccess && !hErr && bad == 0) ? 0 : 1;
}

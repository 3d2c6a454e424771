// tmem_bw.cu -- micro-benchmark: tcgen05.ld throughput / latency of one SM's tensor memory as a
// function of warps per CTA, columns per instruction (x8 / x16 / x32) and loads in flight per wait.
// Used to decide how the tensor-core rollout epilogue should read its accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _variants/tmem_bw tools/tmem_bw.cu && ./_variants/tmem_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ld<8>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

template <int X, int U>
__global__ void bench(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_slot + (uint32_t((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    uint32_t r[U][X];
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < U; ++u) ld<X>(base + uint32_t(((warp >> 2) * U + u) * X) % 512u, r[u]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < U; ++u) {
            asm volatile("" : "+r"(r[u][0]), "+r"(r[u][X - 1])::"memory");
            acc ^= r[u][0] ^ r[u][X - 1];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_slot) : "memory");
}

template <int X, int U>
void run(int warps, long long* d_cycles, uint32_t* d_sink) {
    const int iters = 2000;
    bench<X, U><<<1, warps * 32>>>(200, d_cycles, d_sink);
    bench<X, U><<<1, warps * 32>>>(iters, d_cycles, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("x%d u%d warps %d: %s\n", X, U, warps, cudaGetErrorString(e));
        return;
    }
    long long c;
    cudaMemcpy(&c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
    const double bytes = double(warps) * iters * U * 32.0 * X * 4.0;
    printf("x%-2d in-flight %d warps %2d: %8lld cycles, %7.1f B/clk/SM, %6.1f cycles per wait\n", X, U, warps, c, bytes / double(c),
           double(c) / iters);
}

int main() {
    long long* d_cycles;
    uint32_t* d_sink;
    cudaMalloc(&d_cycles, 8 * 256);
    cudaMalloc(&d_sink, 4);
    for (int warps : {4, 8, 16}) {
        run<8, 1>(warps, d_cycles, d_sink);
        run<8, 2>(warps, d_cycles, d_sink);
        run<8, 4>(warps, d_cycles, d_sink);
        run<16, 1>(warps, d_cycles, d_sink);
        run<16, 2>(warps, d_cycles, d_sink);
        run<16, 4>(warps, d_cycles, d_sink);
        run<32, 1>(warps, d_cycles, d_sink);
        run<32, 2>(warps, d_cycles, d_sink);
    }
    return 0;
}

// Micro-benchmark: how many 32-byte sectors does the memory system move for a sparse 8-byte read?
// Each thread owns one STRIDE-byte record in a buffer far larger than L2 and reads 8 bytes of it, in several ways.
// Finding (B200, CUDA 12.9): see profiles/r01_notes.md "sector accounting".
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dup_sector dup_sector.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
template <int MODE>
__global__ void k(const uint8_t* buf, size_t n, size_t stride, uint32_t off, uint64_t* sink)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if ( t >= n ) return;
    const uint32_t* p = (const uint32_t*)(buf + t * stride + off);
    uint64_t acc = 0;
    __shared__ uint2 sm[256];
    if ( MODE == 0 ) { // plain 8-byte load
        const uint2 v = *(const uint2*)p;
        acc = ((uint64_t)v.x << 32) | v.y;
    }
    else if ( MODE == 1 ) { // two 4-byte loads of the same sector, in flight together
        const uint32_t a = p[0], b = p[1];
        acc = ((uint64_t)a << 32) | b;
    }
    else if ( MODE == 2 ) { // L1 bypass
        const uint2 v = __ldcg((const uint2*)p);
        acc = ((uint64_t)v.x << 32) | v.y;
    }
    else if ( MODE == 3 ) { // read-only path
        const uint2 v = __ldg((const uint2*)p);
        acc = ((uint64_t)v.x << 32) | v.y;
    }
    else if ( MODE == 4 ) { // cp.async 8 bytes (LDGSTS, L1 allocating)
        __pipeline_memcpy_async(&sm[threadIdx.x], p, 8);
        __pipeline_commit();
        __pipeline_wait_prior(0);
        const uint2 v = sm[threadIdx.x];
        acc = ((uint64_t)v.x << 32) | v.y;
    }
    else if ( MODE == 5 ) { // streaming (evict-first) load
        const uint2 v = __ldcs((const uint2*)p);
        acc = ((uint64_t)v.x << 32) | v.y;
    }
    else if ( MODE == 6 ) { // ld.global with an explicit 64-byte L2 prefetch size
        uint2 v;
        asm volatile("ld.global.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
        acc = ((uint64_t)v.x << 32) | v.y;
    }
    else if ( MODE == 7 ) { // volatile
        const uint64_t v = *(const volatile uint64_t*)p;
        acc = v;
    }
    if ( acc == 0x1234567812345678ull ) *sink = acc;
}
template <int MODE>
void run(const uint8_t* buf, size_t n, size_t stride, uint64_t* sink)
{
    k<MODE><<<(unsigned)((n + 255) / 256), 256>>>(buf, n, stride, 8 * 13, sink);
}
int main()
{
    const size_t bytes = 10ull << 30;
    uint8_t* buf; uint64_t* sink;
    if ( cudaMalloc(&buf, bytes) != cudaSuccess ) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes);
    for ( size_t stride : { (size_t)2496, (size_t)128, (size_t)32 } ) {
        const size_t n = stride == 2496 ? (4u << 20) : (16u << 20);
        run<0>(buf, n, stride, sink); run<1>(buf, n, stride, sink); run<2>(buf, n, stride, sink); run<3>(buf, n, stride, sink);
        run<4>(buf, n, stride, sink); run<5>(buf, n, stride, sink); run<6>(buf, n, stride, sink); run<7>(buf, n, stride, sink);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

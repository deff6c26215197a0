// selftest.cu -- TEST INFRASTRUCTURE ONLY: kernels that pin the SIMT interpreter of cuda_emu.h itself (tests/test_emu.py).
// Case 0 checks the collectives, the CTA barrier, atomics and the bulk-copy / mbarrier model against known answers;
// cases 1-4 each contain a deliberate bug the interpreter must abort on (out-of-bounds write, divergent collectives,
// deadlock, a bulk copy nobody waits for); case 5 is a missing __syncwarp that only the reversed lane order exposes.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace
{
__global__ void k_collectives(uint32_t *out, float *fout)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *o = out + warp * 16;
    const uint32_t ballot = __ballot_sync(0xffffffffu, lane % 3 == 0);
    const uint32_t from5 = __shfl_sync(0xffffffffu, (uint32_t)(lane * 7 + warp), 5);
    uint32_t sum = lane + 1;
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    uint32_t inc = 1;
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    const int any = __any_sync(0xffffffffu, lane == 31), all = __all_sync(0xffffffffu, lane < 32), none = __all_sync(0xffffffffu, lane < 31);
    const int mx = __reduce_max_sync(0xffffffffu, lane * (lane & 1 ? -1 : 1)), mn = __reduce_min_sync(0xffffffffu, lane * (lane & 1 ? -1 : 1));
    double dsum = 0.5 * lane;
    for (int d = 16; d > 0; d >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, d);
    __syncwarp();
    if (lane == 7)
    {
        o[0] = ballot; o[1] = from5; o[2] = sum; o[3] = (uint32_t)any; o[4] = (uint32_t)all; o[5] = (uint32_t)none;
        o[6] = (uint32_t)mx; o[7] = (uint32_t)mn; o[8] = (uint32_t)dsum;
    }
    o[9] = 0; // every lane writes the same value
    if (lane == 31) o[10] = inc;
    if (warp == 0 && lane == 0) fout[0] = __fdividef(1.f, 4.f) + rsqrtf(16.f);
}

// exclusive scan of 256 values across the CTA, then a second phase after threads of the upper half have left
__global__ void k_block(uint32_t *out, uint32_t *counter)
{
    __shared__ uint32_t s[256];
    s[threadIdx.x] = threadIdx.x + 1;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1)
    {
        const uint32_t v = threadIdx.x >= (uint32_t)d ? s[threadIdx.x - d] : 0u;
        __syncthreads();
        s[threadIdx.x] += v;
        __syncthreads();
    }
    out[blockIdx.x * 256 + threadIdx.x] = s[threadIdx.x];
    atomicAdd(counter, 1u);
    if (threadIdx.x >= 128) return; // the barrier below only counts the threads that are still alive
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(counter + 1, s[255]);
}

// one warp stages 32 x 16 bytes with a bulk copy and waits on the mbarrier (the pattern of k2_render's CONTIG path)
__global__ void k_bulk(const float4 *src, float4 *dst)
{
    __shared__ __align__(128) float4 buf[32];
    __shared__ __align__(8) unsigned long long bar;
    const int lane = threadIdx.x;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), buf_a = (uint32_t)__cvta_generic_to_shared(buf);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1u) : "memory");
    __syncwarp();
    for (uint32_t phase = 0; phase < 3; ++phase)
    {
        if (lane == 0)
        {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(512u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf_a), "l"(src + 32 * phase), "r"(512u), "r"(bar_a)
                         : "memory");
        }
        uint32_t done = 0;
        while (!done)
        {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(done)
                : "r"(bar_a), "r"(phase & 1u)
                : "memory");
        }
        dst[32 * phase + lane] = buf[lane];
        __syncwarp();
    }
}

__global__ void k_oob(uint32_t *buf, uint32_t n) { buf[n + threadIdx.x] = 1u; }

__global__ void k_divergent(int *out)
{
    if (threadIdx.x & 1) out[0] = __any_sync(0xffffffffu, 1);
    else out[1] = (int)__ballot_sync(0xffffffffu, 1);
}

__global__ void k_deadlock()
{
    if (threadIdx.x == 0) __syncthreads();
    else __syncwarp();
}

__global__ void k_unwaited_copy(const float4 *src)
{
    __shared__ __align__(128) float4 buf[32];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), buf_a = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1u) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(512u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf_a), "l"(src), "r"(512u), "r"(bar_a) : "memory");
    }
}

// a hand-off between lanes through shared memory WITHOUT the __syncwarp it needs: correct only if lane 0 happens to run first
// after the collective (the interpreter continues with the LAST lane to arrive: lane 31 forward, lane 0 reversed)
__global__ void k_missing_syncwarp(uint32_t *out)
{
    __shared__ uint32_t s;
    const int lane = threadIdx.x;
    if (lane == 0) s = 0;
    __syncwarp();
    if (lane == 0) s = 42;
    out[lane] = s; // BUG (deliberate): no __syncwarp() between the write and the reads
}

#define REQUIRE(cond)                                                        \
    do                                                                       \
    {                                                                        \
        if (!(cond))                                                         \
        {                                                                    \
            std::fprintf(stderr, "selftest: %s failed (line %d)\n", #cond, __LINE__); \
            return 1;                                                        \
        }                                                                    \
    } while (0)
} // namespace

extern "C" int emu_selftest(int which)
{
    if (which == 0)
    {
        uint32_t *out = nullptr, *counter = nullptr;
        float *fout = nullptr;
        cudaMalloc((void **)&out, sizeof(uint32_t) * 1024);
        cudaMalloc((void **)&counter, sizeof(uint32_t) * 2);
        cudaMalloc((void **)&fout, sizeof(float) * 4);
        cudaMemsetAsync(counter, 0, sizeof(uint32_t) * 2, nullptr);
        k_collectives<<<2, 64>>>(out, fout);
        std::vector<uint32_t> h(1024);
        cudaMemcpyAsync(h.data(), out, sizeof(uint32_t) * 1024, cudaMemcpyDeviceToHost, nullptr);
        for (int blockwarp = 0; blockwarp < 2; ++blockwarp)
        {
            const uint32_t *o = h.data() + blockwarp * 16;
            REQUIRE(o[0] == 0x49249249u);              // lanes 0, 3, 6, ...
            REQUIRE(o[1] == 35u + (uint32_t)blockwarp); // lane 5's value
            REQUIRE(o[2] == 528u);                     // 1 + ... + 32
            REQUIRE(o[3] == 1u && o[4] == 1u && o[5] == 0u);
            REQUIRE((int)o[6] == 30 && (int)o[7] == -31);
            REQUIRE(o[8] == 248u);                     // 0.5 * (0 + ... + 31)
            REQUIRE(o[9] == 0u && o[10] == 32u);
        }
        float f = 0.f;
        cudaMemcpyAsync(&f, fout, sizeof(float), cudaMemcpyDeviceToHost, nullptr);
        REQUIRE(f == 0.5f);
        k_block<<<3, 256>>>(out, counter);
        cudaMemcpyAsync(h.data(), out, sizeof(uint32_t) * 768, cudaMemcpyDeviceToHost, nullptr);
        for (uint32_t i = 0; i < 768; ++i) REQUIRE(h[i] == ((i % 256) + 1) * ((i % 256) + 2) / 2);
        uint32_t c[2];
        cudaMemcpyAsync(c, counter, sizeof(c), cudaMemcpyDeviceToHost, nullptr);
        REQUIRE(c[0] == 768u && c[1] == 3u * 32896u);
        float4 *src = nullptr, *dst = nullptr;
        cudaMalloc((void **)&src, sizeof(float4) * 96);
        cudaMalloc((void **)&dst, sizeof(float4) * 96);
        std::vector<float> pattern(384);
        for (int i = 0; i < 384; ++i) pattern[i] = (float)i * 0.25f;
        cudaMemcpyAsync(src, pattern.data(), sizeof(float) * 384, cudaMemcpyHostToDevice, nullptr);
        k_bulk<<<1, 32>>>(src, dst);
        std::vector<float> back(384);
        cudaMemcpyAsync(back.data(), dst, sizeof(float) * 384, cudaMemcpyDeviceToHost, nullptr);
        REQUIRE(std::memcmp(back.data(), pattern.data(), sizeof(float) * 384) == 0);
        cudaFree(out); cudaFree(counter); cudaFree(fout); cudaFree(src); cudaFree(dst);
        return 0;
    }
    uint32_t *buf = nullptr;
    cudaMalloc((void **)&buf, sizeof(uint32_t) * 1024);
    if (which == 5)
    {
        // returns how many lanes saw the value: 32 when lane 0 runs first, fewer in the other lane order
        k_missing_syncwarp<<<1, 32>>>(buf);
        uint32_t h[32];
        cudaMemcpyAsync(h, buf, sizeof(h), cudaMemcpyDeviceToHost, nullptr);
        int seen = 0;
        for (uint32_t v : h) seen += v == 42u;
        cudaFree(buf);
        return seen;
    }
    if (which == 1) k_oob<<<1, 32>>>(buf, 1024);
    if (which == 2) k_divergent<<<1, 32>>>((int *)buf);
    if (which == 3) k_deadlock<<<1, 64>>>();
    if (which == 4) k_unwaited_copy<<<1, 32>>>((const float4 *)buf);
    return 2; // not reached for 1..4: the interpreter aborts
}

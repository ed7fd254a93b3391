// Issue-rate probe for the instruction classes of the walk loop on sm_100a: IMAD.WIDE.U32 (Philox mulhilo), LOP3, FSEL, mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe2 tools/probe2.cu ; run on a B200 (gpurun).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
typedef unsigned long long u64;

template <int ILP>
__global__ void k_wide(u32 * out, int iters, u32 m)
{
    u32 lo[ILP], hi[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { lo[j] = threadIdx.x + j; hi[j] = j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            const u64 p = (u64)lo[j]*(u64)m; // IMAD.WIDE.U32
            lo[j] = (u32)p;
            hi[j] ^= (u32)(p >> 32);         // LOP3 (keeps the high half alive)
        }
    }
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s += lo[j] ^ hi[j]; }
    if (s == 0x12345678u) { out[0] = s; }
}

template <int ILP>
__global__ void k_hi(u32 * out, int iters, u32 m)
{
    u32 v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { v[j] = threadIdx.x + j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) { v[j] = __umulhi(v[j], m) + 1u; } // IMAD.HI.U32 with addend
    }
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s += v[j]; }
    if (s == 0x12345678u) { out[0] = s; }
}

template <int ILP>
__global__ void k_lop(u32 * out, int iters, u32 a, u32 b)
{
    u32 v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { v[j] = threadIdx.x + j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) { v[j] = (v[j] ^ a ^ v[(j + 1)%ILP]) ; }
    }
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s += v[j]; }
    if (s == 0x12345678u) { out[0] = s + b; }
}

template <int ILP>
__global__ void k_fsel(float * out, int iters, float a)
{
    float v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { v[j] = threadIdx.x + j; }
    bool p = a > 0.5f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) { v[j] = (v[(j + 1)%ILP] > a) ? v[j] : v[(j + 3)%ILP]; } // FSETP + FSEL
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s += v[j]; }
    if (s == 123.456f || p) { out[0] = s; }
}

// the Philox round mix: 2 IMAD.WIDE + 2 LOP3 per round, ILP independent counters
template <int ILP>
__global__ void k_philox(u32 * out, int iters, u32 k0, u32 k1)
{
    u32 c0[ILP], c1[ILP], c2[ILP], c3[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { c0[j] = threadIdx.x; c1[j] = j; c2[j] = blockIdx.x; c3[j] = 7; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            const u64 p0 = (u64)c0[j]*0xD2511F53ull, p1 = (u64)c2[j]*0xCD9E8D57ull;
            const u32 n0 = (u32)(p1 >> 32) ^ c1[j] ^ k0, n2 = (u32)(p0 >> 32) ^ c3[j] ^ k1;
            c1[j] = (u32)p1; c3[j] = (u32)p0; c0[j] = n0; c2[j] = n2;
        }
    }
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) { s += c0[j] ^ c1[j] ^ c2[j] ^ c3[j]; }
    if (s == 0x12345678u) { out[0] = s; }
}

template <class F>
static double run(F launch, double work)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && work/(ms*1e-3) > best) { best = work/(ms*1e-3); }
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, blocks = sms*8, thr = 256, iters = 20000;
    const double clk = 1.965e9;
    void * d; cudaMalloc(&d, 1024);
    const double base = (double)blocks*thr*iters;
    auto rep = [&](const char * name, double per_s, double inst_per_iter) {
        printf("%-28s %.3e thread-inst/s = %.1f lanes/clk/SM (%.2f cycles per warp instruction per scheduler)\n", name, per_s, per_s/clk/sms, 32.0*4/(per_s/clk/sms));
        (void)inst_per_iter;
    };
    rep("IMAD.WIDE.U32 (+LOP3 each)", run([&] { k_wide<8><<<blocks, thr>>>((u32 *)d, iters, 0xD2511F53u); }, base*8), 2);
    rep("IMAD.HI.U32", run([&] { k_hi<8><<<blocks, thr>>>((u32 *)d, iters, 0xD2511F53u); }, base*8), 1);
    rep("LOP3 (3-input xor)", run([&] { k_lop<8><<<blocks, thr>>>((u32 *)d, iters, 0x9E3779B9u, 1u); }, base*8), 1);
    rep("FSETP+FSEL pair", run([&] { k_fsel<8><<<blocks, thr>>>((float *)d, iters, 0.25f); }, base*8), 2);
    rep("Philox round (2 WIDE+2 LOP3) x4", run([&] { k_philox<4><<<blocks, thr>>>((u32 *)d, iters, 0x9E3779B9u, 0xBB67AE85u); }, base*4), 4);
    rep("Philox round x2 chains", run([&] { k_philox<2><<<blocks, thr>>>((u32 *)d, iters, 0x9E3779B9u, 0xBB67AE85u); }, base*2), 4);
    rep("Philox round x1 chain", run([&] { k_philox<1><<<blocks, thr>>>((u32 *)d, iters, 0x9E3779B9u, 0xBB67AE85u); }, base*1), 4);
    return 0;
}

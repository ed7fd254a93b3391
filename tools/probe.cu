// probe.cu — B200 issue-rate microbenchmarks + prototype Metropolis walk kernel.
// Standalone (nvcc -gencode arch=compute_100a,code=sm_100a probe.cu -o probe -ldl).
// Purpose: (1) prove the NVRTC -> cudaLibraryLoadData -> launch path on the GPU box,
// (2) measure FP64 / INT issue rates (the roofline denominators that MEASURED_PEAKS.json lacks),
// (3) block-size / occupancy sweep of a prototype 3-D Gaussian walk kernel.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <dlfcn.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- microbench
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = threadIdx.x * 1e-9 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = fma(v[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += v[j];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void k_imad(unsigned* out, int iters, unsigned a, unsigned b)
{
    unsigned v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = v[j] * a + b;
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += v[j];
    if (s == 0x12345678u) out[0] = s;
}

template <int ILP>
__global__ void k_lop(unsigned* out, int iters, unsigned a, unsigned b)
{
    unsigned v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) v[j] = (v[j] ^ a) + b; // LOP3 + IADD3 (alu pipe)
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += v[j];
    if (s == 0x12345678u) out[0] = s;
}

// DFMA + IMAD + LOP mixed: can the INT work hide under the FP64 pipe?
template <int ILP>
__global__ void k_mix(double* out, int iters, double a, double b, unsigned ia, unsigned ib)
{
    double v[ILP];
    unsigned w[ILP], z[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) { v[j] = threadIdx.x * 1e-9 + j; w[j] = threadIdx.x + j; z[j] = threadIdx.x * 3 + j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            v[j] = fma(v[j], a, b);
            w[j] = w[j] * ia + ib;
            z[j] = (z[j] ^ ia) + ib;
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += v[j] + w[j] + z[j];
    if (s == 123.456) out[0] = s;
}

__global__ void k_dfma_lat(long long* out, double* dout, int iters, double a, double b)
{
    double v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        v = fma(v, a, b); v = fma(v, a, b); v = fma(v, a, b); v = fma(v, a, b);
        v = fma(v, a, b); v = fma(v, a, b); v = fma(v, a, b); v = fma(v, a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; dout[0] = v; }
}

// ---------------------------------------------------------------- prototype walk
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r > 0) { k.x += W0; k.y += W1; }
        const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    }
    return c;
}

// v in [1,2): 1 + (r+0.5)*2^-32
__device__ __forceinline__ double bits12(unsigned r)
{
    return __hiloint2double(0x3ff00000u | (r >> 12), (r << 20) | 0x80000u);
}

template <int MODE> // 0: full, 1: no exp (a = d), 2: no philox (lcg)
__global__ void k_walk3g(double* __restrict__ x, double step, long long nsteps, unsigned long long seed,
                         unsigned long long offset, double* __restrict__ sums, unsigned* __restrict__ nacc, int W)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double x0 = x[w], x1 = x[W + w], x2 = x[2 * W + w];
    double pold = x0 * x0 + x1 * x1 + x2 * x2;
    double sum = 0.;
    unsigned acc = 0;
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    const double s2 = 2. * step, s3 = -3. * step;
    unsigned lcg = w * 747796405u + 1u;
    for (long long i = 0; i < nsteps; ++i) {
        uint4 r;
        if (MODE == 2) {
            lcg = lcg * 747796405u + 2891336453u; r.x = lcg; lcg = lcg * 747796405u + 2891336453u; r.y = lcg;
            lcg = lcg * 747796405u + 2891336453u; r.z = lcg; lcg = lcg * 747796405u + 2891336453u; r.w = lcg;
        } else {
            const unsigned long long c = offset + (unsigned long long)i;
            r = philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), (unsigned)w, 0u), key);
        }
        const double n0 = fma(s2, bits12(r.x), x0 + s3);
        const double n1 = fma(s2, bits12(r.y), x1 + s3);
        const double n2 = fma(s2, bits12(r.z), x2 + s3);
        const double pn = n0 * n0 + n1 * n1 + n2 * n2;
        const double d = pold - pn;
        const double a = (MODE == 1) ? d + 1. : exp(d);
        const double u = bits12(r.w) - 1.;
        const bool ok = (u <= a);
        x0 = ok ? n0 : x0; x1 = ok ? n1 : x1; x2 = ok ? n2 : x2; pold = ok ? pn : pold;
        acc += ok;
        sum += (x0 * x0 + x1 * x1 + x2 * x2) * (1. / 3.);
    }
    x[w] = x0; x[W + w] = x1; x[2 * W + w] = x2;
    sums[w] = sum; nacc[w] = acc;
}

// ---------------------------------------------------------------- NVRTC path
typedef int (*nvrtcCreateProgram_t)(void**, const char*, const char*, int, const char* const*, const char* const*);
typedef int (*nvrtcCompileProgram_t)(void*, int, const char* const*);
typedef int (*nvrtcGetCUBINSize_t)(void*, size_t*);
typedef int (*nvrtcGetCUBIN_t)(void*, char*);
typedef int (*nvrtcGetProgramLogSize_t)(void*, size_t*);
typedef int (*nvrtcGetProgramLog_t)(void*, char*);
typedef int (*nvrtcDestroyProgram_t)(void**);

static void test_nvrtc()
{
    void* h = dlopen("libnvrtc.so.12", RTLD_NOW);
    if (!h) h = dlopen("/usr/local/cuda/lib64/libnvrtc.so.12", RTLD_NOW);
    if (!h) { printf("NVRTC: dlopen failed: %s\n", dlerror()); return; }
    auto create = (nvrtcCreateProgram_t)dlsym(h, "nvrtcCreateProgram");
    auto compile = (nvrtcCompileProgram_t)dlsym(h, "nvrtcCompileProgram");
    auto cubsz = (nvrtcGetCUBINSize_t)dlsym(h, "nvrtcGetCUBINSize");
    auto cub = (nvrtcGetCUBIN_t)dlsym(h, "nvrtcGetCUBIN");
    auto logsz = (nvrtcGetProgramLogSize_t)dlsym(h, "nvrtcGetProgramLogSize");
    auto logget = (nvrtcGetProgramLog_t)dlsym(h, "nvrtcGetProgramLog");
    auto destroy = (nvrtcDestroyProgram_t)dlsym(h, "nvrtcDestroyProgram");
    const char* src = "extern \"C\" __global__ void jit_k(double* o, double a) { o[threadIdx.x] = exp(a * threadIdx.x); }\n";
    void* prog = nullptr;
    int rc = create(&prog, src, "jit.cu", 0, nullptr, nullptr);
    const char* opts[] = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17"};
    rc = compile(prog, 3, opts);
    size_t ls = 0; logsz(prog, &ls);
    if (ls > 1) { std::string log(ls, 0); logget(prog, &log[0]); printf("NVRTC log: %s\n", log.c_str()); }
    if (rc != 0) { printf("NVRTC: compile failed rc=%d\n", rc); return; }
    size_t cs = 0; cubsz(prog, &cs);
    std::vector<char> cubin(cs); cub(prog, cubin.data()); destroy(&prog);
    cudaLibrary_t lib; cudaKernel_t kern;
    CK(cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    CK(cudaLibraryGetKernel(&kern, lib, "jit_k"));
    double* d; CK(cudaMalloc(&d, 32 * 8));
    double a = 0.5; void* args[] = {&d, &a};
    CK(cudaLaunchKernel((const void*)kern, dim3(1), dim3(32), args, 0, 0));
    CK(cudaDeviceSynchronize());
    double hbuf[32]; CK(cudaMemcpy(hbuf, d, sizeof(hbuf), cudaMemcpyDeviceToHost));
    printf("NVRTC: cubin %zu bytes, jit_k[2]=%.17g (expect e=2.7182818284590451) -> %s\n", cs, hbuf[2],
           hbuf[2] == 2.7182818284590451 ? "OK" : "MISMATCH");
    cudaFree(d);
}

template <typename F>
static float time_ms(F f, int reps = 3)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv)
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s sm_%d%d SMs=%d clockRate=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, clk);
    const int SM = p.multiProcessorCount;
    test_nvrtc();

    double* dout; CK(cudaMalloc(&dout, 1024)); unsigned* uout = (unsigned*)dout;
    // --- throughput: grid = SM*8 blocks x 256 threads (full occupancy 2048 thr/SM)
    {
        const int iters = 20000; const int blocks = SM * 8, thr = 256;
        const double nthreads = (double)blocks * thr;
        float ms;
        ms = time_ms([&] { k_dfma<8><<<blocks, thr>>>(dout, iters, 1.0000001, 1e-9); });
        double dfma_rate = nthreads * iters * 8 / (ms * 1e-3);
        printf("DFMA  : %.3f ms  %.4e DFMA/s  = %.2f lanes/SM/clk@1965MHz  (%.2f TFLOP/s)\n", ms, dfma_rate, dfma_rate / SM / 1.965e9, 2 * dfma_rate / 1e12);
        ms = time_ms([&] { k_imad<8><<<blocks, thr>>>(uout, iters, 747796405u, 12345u); });
        double r = nthreads * iters * 8 / (ms * 1e-3);
        printf("IMAD  : %.3f ms  %.4e IMAD/s  = %.2f lanes/SM/clk@1965MHz\n", ms, r, r / SM / 1.965e9);
        ms = time_ms([&] { k_lop<8><<<blocks, thr>>>(uout, iters, 747796405u, 12345u); });
        r = nthreads * iters * 8 * 2 / (ms * 1e-3);
        printf("LOP+IADD: %.3f ms  %.4e ops/s  = %.2f lanes/SM/clk@1965MHz\n", ms, r, r / SM / 1.965e9);
        ms = time_ms([&] { k_mix<4><<<blocks, thr>>>(dout, iters, 1.0000001, 1e-9, 747796405u, 12345u); });
        r = nthreads * iters * 4 / (ms * 1e-3);
        printf("MIX (1 DFMA + 1 IMAD + LOP + IADD per iter): %.3f ms  %.4e iter/s = %.2f DFMA lanes/SM/clk (x4 total instr)\n", ms, r, r / SM / 1.965e9);
        long long* lat; CK(cudaMalloc(&lat, 8));
        k_dfma_lat<<<1, 32>>>(lat, dout, 1000, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        long long hl; CK(cudaMemcpy(&hl, lat, 8, cudaMemcpyDeviceToHost));
        printf("DFMA dependent latency: %.2f cycles\n", hl / 8000.0);
    }
    // --- prototype walk kernel sweeps
    {
        const long long nsteps = argc > 1 ? atoll(argv[1]) : 20000;
        int Ws[] = {65536, 148 * 512, 148 * 1024, 148 * 2048, 148 * 4096};
        for (int W : Ws) {
            double* x; double* sums; unsigned* nacc;
            CK(cudaMalloc(&x, 3 * W * 8)); CK(cudaMalloc(&sums, W * 8)); CK(cudaMalloc(&nacc, W * 4));
            int bss[] = {32, 64, 128, 256, 512};
            for (int bs : bss) {
                CK(cudaMemset(x, 0, 3 * W * 8));
                const int blocks = (W + bs - 1) / bs;
                float ms = time_ms([&] { k_walk3g<0><<<blocks, bs>>>(x, 1.0, nsteps, 1337ull, 0ull, sums, nacc, W); }, 2);
                std::vector<double> hs(W); std::vector<unsigned> ha(W);
                CK(cudaMemcpy(hs.data(), sums, W * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ha.data(), nacc, W * 4, cudaMemcpyDeviceToHost));
                double tot = 0, ta = 0; for (int i = 0; i < W; ++i) { tot += hs[i]; ta += ha[i]; }
                printf("walk3g W=%7d bs=%3d: %8.3f ms  %.4e steps/s   <x^2>=%.6f acc=%.5f\n", W, bs, ms, (double)W * nsteps / (ms * 1e-3),
                       tot / ((double)W * nsteps), ta / ((double)W * nsteps));
            }
            // ablations at bs=64
            {
                const int bs = 64, blocks = (W + bs - 1) / bs;
                CK(cudaMemset(x, 0, 3 * W * 8));
                float ms1 = time_ms([&] { k_walk3g<1><<<blocks, bs>>>(x, 1.0, nsteps, 1337ull, 0ull, sums, nacc, W); }, 2);
                CK(cudaMemset(x, 0, 3 * W * 8));
                float ms2 = time_ms([&] { k_walk3g<2><<<blocks, bs>>>(x, 1.0, nsteps, 1337ull, 0ull, sums, nacc, W); }, 2);
                printf("   ablation W=%d bs=64: no-exp %.3f ms (%.4e/s), lcg-rng %.3f ms (%.4e/s)\n", W, ms1, (double)W * nsteps / (ms1 * 1e-3), ms2,
                       (double)W * nsteps / (ms2 * 1e-3));
            }
            cudaFree(x); cudaFree(sums); cudaFree(nacc);
        }
    }
    return 0;
}

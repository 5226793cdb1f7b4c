// Microbenchmark: issue rate of FFMA2 (fma.rn.f32x2) by operand form, B200.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o ffma2_forms ffma2_forms.cu && ./ffma2_forms
// Reports cycles per FFMA2 per SM sub-partition (warp scheduler); 2.0 = the pipe's nominal rate.
#include <cstdio>
#include <cuda_runtime.h>

#define N_ACC 8
template <int FORM>
__global__ void __launch_bounds__(640, 1) k(float2 *out, int iters, float2 a, float2 b, float c, long long *cyc) {
    float2 x[N_ACC], y[N_ACC];
    for (int i = 0; i < N_ACC; ++i) { x[i] = make_float2(threadIdx.x + i, 1.f + i); y[i] = make_float2(0.5f * i, 2.f); }
    c += (float)threadIdx.x * 1e-9f;          // a per-thread value: an ordinary register, not a uniform one
    a.x += c * 1e-9f; b.y += c * 1e-9f;
    const float2 cc = make_float2(c, c);      // both halves equal: the compiler emits the scalar-broadcast form (.F32)
    const float2 nc = make_float2(-c, -c);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N_ACC; ++i) {
            if (FORM == 0) x[i] = __ffma2_rn(x[i], a, b);               // packed * packed + packed (accumulator chain)
            if (FORM == 1) x[i] = __ffma2_rn(x[i], cc, b);              // packed * broadcast scalar + packed
            if (FORM == 2) x[i] = __ffma2_rn(y[i], cc, x[i]);           // like the cull chain: fresh pair * scalar + accumulator
            if (FORM == 3) x[i] = __ffma2_rn(y[i], nc, x[i]);           // ... with a negated scalar
            if (FORM == 4) x[i] = __ffma2_rn(x[i], x[i], y[i]);         // s * s + nq
            if (FORM == 5) { x[i] = __ffma2_rn(y[i], cc, x[i]); y[i] = __ffma2_rn(x[i], a, y[i]); }   // two dependent per acc
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < N_ACC; ++i) s += x[i].x + x[i].y + y[i].x;
    if (s == 123456.f) out[0] = make_float2(s, s);
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int FORM>
static void run(const char *what, int per_iter) {
    float2 *out; long long *cyc, h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<FORM><<<148, 640>>>(out, iters, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, -1e-3f), 1.00001f, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<FORM><<<148, 640>>>(out, iters, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, -1e-3f), 1.00001f, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_per_sched = 640 / 32 / 4.0;
    const double n = (double)iters * per_iter * warps_per_sched;      // FFMA2 per scheduler
    printf("%-58s %.3f cycles / FFMA2 / scheduler   (%.1f TFLOP/s, %.3f ms)\n", what, (double)h / n,
           148.0 * 640 * iters * per_iter * 4.0 / (ms * 1e-3) / 1e12, ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("x = x * a + b           (all packed)", N_ACC);
    run<1>("x = x * (c,c) + b       (scalar broadcast multiplier)", N_ACC);
    run<2>("x = y * (c,c) + x       (cull chain form)", N_ACC);
    run<3>("x = y * (-c,-c) + x     (negated scalar)", N_ACC);
    run<4>("x = x * x + y", N_ACC);
    run<5>("x = y*(c,c)+x ; y = x*a+y", 2 * N_ACC);
    return 0;
}

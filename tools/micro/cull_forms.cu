// Microbenchmark: the FFMA2 stream of the brute-force cull (rg_trace.cuh cull_h2), by itself, in several arrangements.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o cull_forms cull_forms.cu && ./cull_forms
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float max3nan(float a, float b, float c) { float d; asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float max2nan(float a, float b) { float d; asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

struct CR { float2 dx, dy, dz, nod, px, py, pz; float nthr; };

template <int FORM, int R>
__global__ void __launch_bounds__(640, 1) k(const float4 *recs, int n, float *out, long long *cyc, float seed) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = recs[i];
    __syncthreads();
    CR cr[R];
    for (int r = 0; r < R; ++r) {
        float v = seed + threadIdx.x * 1e-3f + r;
        if (FORM == 1) {   // genuinely different halves: the compiler must keep pairs (no .F32 broadcast form)
            cr[r].dx = make_float2(v, v * 1.0000001f); cr[r].dy = make_float2(v * 2, v * 2.0000001f); cr[r].dz = make_float2(v * 3, v * 3.0000001f);
            cr[r].nod = make_float2(v * 4, v * 4.0000001f); cr[r].px = make_float2(v * 5, v * 5.0000001f); cr[r].py = make_float2(v * 6, v * 6.0000001f);
            cr[r].pz = make_float2(v * 7, v * 7.0000001f);
        } else {
            cr[r].dx = make_float2(v, v); cr[r].dy = make_float2(v * 2, v * 2); cr[r].dz = make_float2(v * 3, v * 3);
            cr[r].nod = make_float2(v * 4, v * 4); cr[r].px = make_float2(v * 5, v * 5); cr[r].py = make_float2(v * 6, v * 6);
            cr[r].pz = make_float2(v * 7, v * 7);
        }
        cr[r].nthr = -1e30f * seed;
    }
    float acc = 0.f;
    bool any = false;
    float2 sum = make_float2(0.f, 0.f);
    long long t0 = clock64();
    if (FORM >= 6) {
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll 1
            for (int j = 0; j < n; j += 4) {
                const float4 A0 = sm[j], B0 = sm[j + 1], A1 = sm[j + 2], B1 = sm[j + 3];
                float m[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float2 h0, h1;
                    {
                        const float2 X = make_float2(A0.x, A0.y), Y = make_float2(A0.z, A0.w), Z = make_float2(B0.x, B0.y), NK = make_float2(B0.z, B0.w);
                        const float2 s_ = __ffma2_rn(X, cr[r].dx, __ffma2_rn(Y, cr[r].dy, __ffma2_rn(Z, cr[r].dz, cr[r].nod)));
                        const float2 nq = __ffma2_rn(X, cr[r].px, __ffma2_rn(Y, cr[r].py, __ffma2_rn(Z, cr[r].pz, NK)));
                        h0 = __ffma2_rn(s_, s_, nq);
                    }
                    {
                        const float2 X = make_float2(A1.x, A1.y), Y = make_float2(A1.z, A1.w), Z = make_float2(B1.x, B1.y), NK = make_float2(B1.z, B1.w);
                        const float2 s_ = __ffma2_rn(X, cr[r].dx, __ffma2_rn(Y, cr[r].dy, __ffma2_rn(Z, cr[r].dz, cr[r].nod)));
                        const float2 nq = __ffma2_rn(X, cr[r].px, __ffma2_rn(Y, cr[r].py, __ffma2_rn(Z, cr[r].pz, NK)));
                        h1 = __ffma2_rn(s_, s_, nq);
                    }
                    if (FORM == 6) { any = any | !(max2nan(h0.x, h0.y) < cr[r].nthr) | !(max2nan(h1.x, h1.y) < cr[r].nthr); }
                    if (FORM == 7) { m[r] = max2nan(max3nan(h0.x, h0.y, h1.x), h1.y); any = any | !(m[r] < cr[r].nthr); }
                    if (FORM == 8) { m[r] = max2nan(max3nan(h0.x, h0.y, h1.x), h1.y) - cr[r].nthr; }
                }
                if (FORM == 8) { float mm = m[0]; for (int r = 1; r < R; ++r) mm = max2nan(mm, m[r]); any = any | !(mm < 0.f); }
            }
        }
    } else
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll 1
        for (int j = 0; j < n; j += 2) {
            const float4 A = sm[j], B = sm[j + 1];
            const float2 X = make_float2(A.x, A.y), Y = make_float2(A.z, A.w), Z = make_float2(B.x, B.y), NK = make_float2(B.z, B.w);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 h;
                if (FORM == 2) {   // Horner-like single chain order (z first), as the library does
                    const float2 s = __ffma2_rn(X, cr[r].dx, __ffma2_rn(Y, cr[r].dy, __ffma2_rn(Z, cr[r].dz, cr[r].nod)));
                    const float2 nq = __ffma2_rn(X, cr[r].px, __ffma2_rn(Y, cr[r].py, __ffma2_rn(Z, cr[r].pz, NK)));
                    h = __ffma2_rn(s, s, nq);
                } else {
                    const float2 s = __ffma2_rn(X, cr[r].dx, __ffma2_rn(Y, cr[r].dy, __ffma2_rn(Z, cr[r].dz, cr[r].nod)));
                    const float2 nq = __ffma2_rn(X, cr[r].px, __ffma2_rn(Y, cr[r].py, __ffma2_rn(Z, cr[r].pz, NK)));
                    h = __ffma2_rn(s, s, nq);
                }
                if (FORM == 3) any = any | !(h.x < cr[r].nthr);                        // half the compares (measurement only)
                else if (FORM == 4) sum = __ffma2_rn(h, cr[r].dx, sum);                 // no compares at all: one more FFMA2 instead
                else if (FORM == 5) any = any | !(fmaxf(h.x, h.y) < cr[r].nthr);       // max + one compare
                else any = any | !(h.x < cr[r].nthr) | !(h.y < cr[r].nthr);
            }
        }
    }
    long long t1 = clock64();
    if (any) acc = 1.f;
    if (sum.x + sum.y == 77.f) acc = 2.f;
    if (acc != 0.f) out[threadIdx.x & 1] = acc;   // (never taken: every h is far above the threshold... but the compiler cannot know)
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int FORM, int R>
static void run(const char *what) {
    const int n = 8192;
    float4 *recs; float *out; long long *cyc, h;
    cudaMalloc(&recs, n * 16 + 64); cudaMemset(recs, 0, n * 16 + 64); cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(k<FORM, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 16);
    for (int i = 0; i < 2; ++i) k<FORM, R><<<148, 640, n * 16>>>(recs, n, out, cyc, 0.5f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double groups = 8.0 * (n / (FORM >= 6 ? 4 : 2)) * (640 / 32 / 4.0);     // warp-groups per scheduler
    printf("%-64s %.2f cycles per group of %d FFMA2 = %.3f cycles / FFMA2   (%s)\n", what, (double)h / groups, (FORM == 4 ? 8 : FORM >= 6 ? 14 : 7) * R,
           (double)h / groups / ((FORM == 4 ? 8 : FORM >= 6 ? 14 : 7) * R), cudaGetErrorString(cudaGetLastError()));
    cudaFree(recs); cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 3>("R=3, ray constants with equal halves (scalar-broadcast form)");
    run<1, 3>("R=3, ray constants as true pairs");
    run<0, 2>("R=2, scalar-broadcast form");
    run<1, 2>("R=2, true pairs");
    run<0, 4>("R=4, scalar-broadcast form");
    run<3, 3>("R=3, 3 compares instead of 6");
    run<4, 3>("R=3, no compares, 8 FFMA2 per ray (24 per group)");
    run<5, 3>("R=3, max + 1 compare per ray");
    run<6, 3>("R=3, 2 groups per turn: max.NaN + compare per ray per group (cycles per 2 groups)");
    run<7, 3>("R=3, 2 groups per turn: max3.NaN, max.NaN, 1 compare per ray (per 2 groups)");
    run<8, 3>("R=3, 2 groups per turn: ... minus threshold, max over rays, 1 compare (per 2 groups)");
    return 0;
}

// Micro-benchmark (tooling, not product): FP32 issue rates on sm_100a that decide the
// channelizer's inner-loop shape: scalar FMUL+FADD vs FFMA vs packed f32x2 (FMUL2/FADD2/FFMA2),
// with register, shared-broadcast and __constant__ second operands.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

__constant__ float4 cw[256];

template<int MODE> __global__ void __launch_bounds__(256) k(float *out, const float4 *gw, int iters)
{
    __shared__ float4 sw[256];
    sw[threadIdx.x] = gw[threadIdx.x];
    __syncthreads();
    float x = threadIdx.x * 0.001f + 1.0f, y = 0.5f + blockIdx.x * 1e-3f;
    float2 acc[8];
    #pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; it++) {
        #pragma unroll 8
        for (int j = 0; j < 32; j++) {
            float a = x + j, b = y - j;   // stand-in for the converted sample
            #pragma unroll
            for (int c = 0; c < 8; c++) {
                float4 w;
                if (MODE >= 10 && MODE < 20) w = sw[(j * 8 + c) & 255];
                else if (MODE >= 20) w = cw[(j * 8 + c) & 255];
                else w = make_float4(x + c, y + c, -(y + c), x + c);
                int m = MODE % 10;
                if (m == 0) {        // exact scalar: 4 FMUL + 4 FADD
                    float pr = __fadd_rn(__fmul_rn(a, w.x), __fmul_rn(b, w.z));
                    float pi = __fadd_rn(__fmul_rn(a, w.y), __fmul_rn(b, w.w));
                    acc[c].x = __fadd_rn(acc[c].x, pr);
                    acc[c].y = __fadd_rn(acc[c].y, pi);
                } else if (m == 1) { // scalar FFMA: 4 FFMA
                    acc[c].x = fmaf(a, w.x, acc[c].x); acc[c].x = fmaf(b, w.z, acc[c].x);
                    acc[c].y = fmaf(a, w.y, acc[c].y); acc[c].y = fmaf(b, w.w, acc[c].y);
                } else if (m == 2) { // exact packed: 2 FMUL2 + 2 FADD2
                    float2 p1 = __fmul2_rn(make_float2(a, a), make_float2(w.x, w.y));
                    float2 p2 = __fmul2_rn(make_float2(b, b), make_float2(w.z, w.w));
                    acc[c] = __fadd2_rn(acc[c], __fadd2_rn(p1, p2));
                } else if (m == 3) { // packed FFMA2: 2
                    acc[c] = __ffma2_rn(make_float2(a, a), make_float2(w.x, w.y), acc[c]);
                    acc[c] = __ffma2_rn(make_float2(b, b), make_float2(w.z, w.w), acc[c]);
                } else if (m == 4) { // exact, FMUL2 for products, scalar adds
                    float2 p1 = __fmul2_rn(make_float2(a, a), make_float2(w.x, w.y));
                    float2 p2 = __fmul2_rn(make_float2(b, b), make_float2(w.z, w.w));
                    acc[c].x = __fadd_rn(acc[c].x, __fadd_rn(p1.x, p2.x));
                    acc[c].y = __fadd_rn(acc[c].y, __fadd_rn(p1.y, p2.y));
                }
            }
        }
        x += 1e-7f;
    }
    float s = 0;
    #pragma unroll
    for (int c = 0; c < 8; c++) s += acc[c].x + acc[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int MODE> int run(const char *name, float *out, float4 *gw, int blocks)
{
    int iters = 200;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, gw, 10);
    CK(cudaDeviceSynchronize());
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, gw, iters);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double macs = (double)blocks * 256 * iters * 32 * 8;   // complex MACs
    printf("%-34s %8.3f ms  %8.1f G complex-MAC/s  (%.1f per clk per SM @1.9GHz)\n", name, best, macs / best * 1e-6,
           macs / best * 1e-6 / 148 / 1.9);
    return 0;
}

int main()
{
    float *out; float4 *gw; float4 h[256];
    for (int i = 0; i < 256; i++) h[i] = make_float4(1.f + i * 1e-3f, 0.5f, -0.5f, 1.f + i * 1e-3f);
    CK(cudaMalloc(&out, 148 * 16 * 256 * 4)); CK(cudaMalloc(&gw, sizeof(h)));
    CK(cudaMemcpy(gw, h, sizeof(h), cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(cw, h, sizeof(h)));
    int blocks = 148 * 8;
    run<0>("reg  exact scalar 4FMUL+4FADD", out, gw, blocks);
    run<1>("reg  scalar 4FFMA", out, gw, blocks);
    run<2>("reg  exact packed 2FMUL2+2FADD2", out, gw, blocks);
    run<3>("reg  packed 2FFMA2", out, gw, blocks);
    run<4>("reg  exact 2FMUL2+4FADD", out, gw, blocks);
    run<10>("smem exact scalar", out, gw, blocks);
    run<11>("smem scalar FFMA", out, gw, blocks);
    run<12>("smem exact packed", out, gw, blocks);
    run<13>("smem packed FFMA2", out, gw, blocks);
    run<20>("cmem exact scalar", out, gw, blocks);
    run<21>("cmem scalar FFMA", out, gw, blocks);
    run<22>("cmem exact packed", out, gw, blocks);
    run<23>("cmem packed FFMA2", out, gw, blocks);
    return 0;
}

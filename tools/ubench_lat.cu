// Latency micro-benchmark (tooling): dependent-chain cycles per op for the FP64 / conversion ops on
// the demod kernel's critical path, one warp, clock64 timing.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template<int OP> __global__ void k(double *out, double a, double b, long long *cyc)
{
    double x = a + threadIdx.x * 1e-9, y = b; float f = (float)a; int acc = 0;
    long long t0 = clock64();
    #pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = __dadd_rn(x, y);
        if (OP == 1) x = fma(x, y, y);
        if (OP == 2) x = __dmul_rn(x, y);
        if (OP == 3) { f = __double2float_rn(x); x = (double)f + y; }          // F2F both ways + DADD
        if (OP == 4) x = __ddiv_rn(x, y) + 1.0;
        if (OP == 5) x = __dsqrt_rn(x) + 3.0;
        if (OP == 6) { double s, c; sincos(x, &s, &c); x = s + c + 1.0; }
        if (OP == 7) { f = __fadd_rn(f, 1.0f); }
        if (OP == 8) { f = __fmul_rn(f, 1.0001f); }
        if (OP == 9) { x = (x >= y) ? __dadd_rn(x, -y) : x; x = __dadd_rn(x, 0.9); }   // wrap step
        if (OP == 10) { unsigned long long u = __double_as_longlong(__dadd_rn(x, y)); u += 0x0FFFFFFFull + ((u >> 29) & 1ull); u &= ~0x1FFFFFFFull; x = __longlong_as_double(u); }
        if (OP == 11) { acc += __double2int_rz(x); x = __dadd_rn(x, (double)(acc & 1)); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + f + acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template<int OP> void run(const char *name, double a, double b)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    k<OP><<<1, 32>>>(out, a, b, cyc); cudaDeviceSynchronize();
    k<OP><<<1, 32>>>(out, a, b, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %7.1f cycles/iter\n", name, (double)h / N);
}
int main()
{
    run<0>("DADD chain", 1.0, 1e-9);
    run<1>("DFMA chain", 1.0, 0.5);
    run<2>("DMUL chain", 1.0, 1.0000001);
    run<3>("F2F.f32<-f64 + F2F back + DADD", 1.0, 0.3);
    run<4>("ddiv_rn + DADD", 3.0, 1.7);
    run<5>("dsqrt_rn + DADD", 5.0, 1.0);
    run<6>("sincos(double) + 2 DADD", 0.7, 1.0);
    run<7>("FADD chain", 1.0, 1.0);
    run<8>("FMUL chain", 1.0, 1.0);
    run<9>("phase wrap step (DSETP+DADD+sel+DADD)", 1.0, 6.283);
    run<10>("DADD + integer round_to_f32", 1.0, 0.9);
    run<11>("D2I + I2D + DADD", 1.0, 0.9);
    return 0;
}

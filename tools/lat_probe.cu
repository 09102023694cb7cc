// Dependent-chain latencies on the device the SGS dataflow kernel cares about: fp64 mul/add/div, 64-bit shuffle, LDS, L2 load, flag round trip.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, long long *cyc, const double *g, int n)
{
	__shared__ double s[64];
	const int lane = threadIdx.x;
	s[lane] = 1.0 + lane * 1e-9; s[lane + 32] = 1.0;
	__syncwarp();
	double a = 1.0 + lane * 1e-12, b = 1.0000001;
	long long t0 = clock64();
	for (int i = 0; i < n; i++) a = __dmul_rn(a, b);
	long long t1 = clock64();
	for (int i = 0; i < n; i++) a = __dadd_rn(a, b);
	long long t2 = clock64();
	for (int i = 0; i < n; i++) a = __ddiv_rn(b, a);
	long long t3 = clock64();
	for (int i = 0; i < n; i++) a = __shfl_up_sync(0xffffffffu, a, 1);
	long long t4 = clock64();
	int idx = lane;
	for (int i = 0; i < n; i++) idx = (int)s[idx] + lane - 1;   // LDS.64 + cvt chain
	long long t5 = clock64();
	const double *p = g + lane;
	double acc = 0;
	for (int i = 0; i < n; i++) { double v = __ldcg(p + (int)acc); acc = v * 0.0; }   // dependent L2 loads (values are zero)
	long long t6 = clock64();
	for (int i = 0; i < n; i++) a = fma(a, b, a);
	long long t7 = clock64();
	out[lane] = a + idx + acc;
	if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; }
}
// flag ping-pong between two SMs through L2: round trips
__global__ void pingpong(volatile int *flag, int n, long long *cyc)
{
	if (threadIdx.x) return;
	const int me = blockIdx.x;
	long long t0 = clock64();
	for (int i = 0; i < n; i++) {
		if (me == 0) { *flag = 2 * i + 1; while (*flag != 2 * i + 2) { } }
		else { while (*flag != 2 * i + 1) { } *flag = 2 * i + 2; }
	}
	if (me == 0) cyc[0] = clock64() - t0;
}
__global__ void pingpong_fence(volatile int *flag, int n, long long *cyc, double *data)
{
	if (threadIdx.x) return;
	const int me = blockIdx.x;
	long long t0 = clock64();
	for (int i = 0; i < n; i++) {
		if (me == 0) { data[i & 1023] = i; __threadfence(); *flag = 2 * i + 1; while (*flag != 2 * i + 2) { } __threadfence(); }
		else { while (*flag != 2 * i + 1) { } __threadfence(); data[1024 + (i & 1023)] = i; __threadfence(); *flag = 2 * i + 2; }
	}
	if (me == 0) cyc[0] = clock64() - t0;
}
int main()
{
	double *out, *g; long long *cyc; int *flag;
	cudaMalloc(&out, 4096 * 8); cudaMalloc(&g, 1 << 20); cudaMemset(g, 0, 1 << 20); cudaMalloc(&cyc, 64); cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4);
	const int n = 4096;
	lat<<<1, 32>>>(out, cyc, g, n); lat<<<1, 32>>>(out, cyc, g, n);
	long long h[8]; cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
	const char *nm[] = {"dmul", "dadd", "ddiv", "shfl64", "lds64+cvt", "ldcg(L2)", "dfma"};
	for (int i = 0; i < 7; i++) printf("%-10s %8.1f cycles per dependent op\n", nm[i], (double)h[i] / n);
	pingpong<<<2, 32>>>(flag, 2000, cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
	printf("flag ping-pong (volatile, 2 SMs) %8.1f cycles per round trip\n", (double)h[0] / 2000);
	cudaMemset(flag, 0, 4);
	pingpong_fence<<<2, 32>>>(flag, 2000, cyc, out); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
	printf("flag ping-pong with store+__threadfence each side %8.1f cycles per round trip\n", (double)h[0] / 2000);
	int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz; %s\n", clk, cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}

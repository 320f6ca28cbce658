/* k_peak.cu -- measured denominator of the FP32-issue roofline ("measure it with an FMUL/FADD micro-benchmark
 * in the same harness", because the box's sustained SM clock need not be the nominal one).
 *
 * The traversal kernels (k_trace, k_shadow) are bound by FP32 instruction issue, and the library is compiled with
 * -fmad=false (parity forbids FMA contraction), so their arithmetic is FMUL + FADD, one lane-op each.  This kernel issues
 * exactly that mix from eight independent chains per thread with every SM fully occupied; lane-ops per second of it is
 * the peak the `fp32_issue` fractions of bench.py are quoted against (nominal: SMs x 128 lanes x clock).
 */
#include "k_common.cuh"

#define PK_CHAINS 8
#define PK_ITERS 4096
#define PK_THREADS 256

__global__ void __launch_bounds__(PK_THREADS) k_fp32_issue(float *out, float m, float c)
{
	float x[PK_CHAINS];
#pragma unroll
	for (int k = 0; k < PK_CHAINS; k++) x[k] = (float)(threadIdx.x + k) * 1.0e-3f;
#pragma unroll 1
	for (int it = 0; it < PK_ITERS / 8; it++) {
#pragma unroll
		for (int u = 0; u < 8; u++)
#pragma unroll
			for (int k = 0; k < PK_CHAINS; k++) x[k] = x[k] * m + c; /* FMUL, FADD (no contraction) */
	}
	float s = 0.0f;
#pragma unroll
	for (int k = 0; k < PK_CHAINS; k++) s += x[k];
	if (s == 123456.789f) out[0] = s; /* keeps the chains alive; practically never true */
}

/* lane-ops (FMUL + FADD) per second, best of `reps` launches, timed with CUDA events on the context's stream */
extern "C" int ug_fp32_issue_peak(ug_context *ctx, double *lane_ops_per_s)
{
	if (!ctx || !lane_ops_per_s) return UG_ERR_ARG;
	float *d = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	UG_CUDA(ctx, cudaMalloc(&d, 256));
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int blocks = ctx->num_sms * (2048 / PK_THREADS);
	const double ops = (double)blocks * PK_THREADS * (double)PK_ITERS * PK_CHAINS * 2.0;
	double best = 0.0;
	cudaError_t err = cudaSuccess;
	for (int rep = 0; rep < 6 && err == cudaSuccess; rep++) { /* the first launches warm the clocks up */
		cudaEventRecord(e0, ctx->stream);
		k_fp32_issue<<<blocks, PK_THREADS, 0, ctx->stream>>>(d, 0.999f, 1.0e-3f);
		cudaEventRecord(e1, ctx->stream);
		err = cudaEventSynchronize(e1);
		float ms = 0.0f;
		if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
		if (err == cudaSuccess && ms > 0.0f && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	cudaFree(d);
	if (err != cudaSuccess) return ug_fail(ctx, UG_ERR_CUDA, "ug_fp32_issue_peak: %s", cudaGetErrorString(err));
	*lane_ops_per_s = best;
	return UG_OK;
}

/* k_sort.cu -- in-house scan and stable LSD radix sort, replacing every CUDPP call of the reference
 * (cudpp/cudpp.h:427-471; call sites frustum_grid.h:249,298,334,361 and decision_data.h:177,209,261).
 *
 * All sizes that depend on the scene (number of (cell,triangle) pairs) are read from device memory,
 * so the launches have fixed shapes and the whole frame can be replayed as one CUDA graph.
 *
 * Radix sort = classic reduce-then-scan LSD, 8-bit digits, fixed grid of G persistent blocks:
 *   pass:  rs_hist     block b counts the digits of its contiguous chunk      -> hist[d][b]
 *          rs_scan     exclusive scan of hist in digit-major order            -> base[d][b]
 *          rs_scatter  block b re-reads its chunk in order, ranks each key among equal digits with
 *                      warp match-any (order = warp, round, lane = input order) and writes it to
 *                      base[d][b] + rank   => stable, which is what gives ascending triangle ids
 *                      inside a cell (frustum_grid.h:298 relies on cudppSort being stable).
 * Memory-bound: 2 reads + 1 write of 8 B per pair and pass.
 */
#include "k_grid_dev.cuh"

#ifndef RS_BLOCKS_PER_SM
#define RS_BLOCKS_PER_SM 2
#endif

__global__ void __launch_bounds__(RS_THREADS) rs_hist(const u32 *__restrict__ keys, const u32 *__restrict__ n_dev, u32 n_max, int shift,
                                                      u32 *__restrict__ hist)
{
	__shared__ u32 s_hist[257];
	const u32 n = min(*n_dev, n_max), G = gridDim.x, b = blockIdx.x;
	s_hist[threadIdx.x] = 0;
	__syncthreads();
	u32 lo, hi;
	rs_chunk(n, G, b, lo, hi);
	rs_block_hist<RS_THREADS>(keys, lo, hi, shift, s_hist);
	__syncthreads();
	hist[(size_t)threadIdx.x * G + b] = s_hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const u32 *__restrict__ keys_in, const u32 *__restrict__ vals_in, u32 *__restrict__ keys_out,
                                                         u32 *__restrict__ vals_out, const u32 *__restrict__ n_dev, u32 n_max, int shift,
                                                         const u32 *__restrict__ base)
{
	__shared__ u32 s_cnt[RS_THREADS / 32][257]; /* per warp, per digit; [256] = dummy bin of out-of-range lanes */
	__shared__ u32 s_run[256];           /* where the block's next key of each digit goes */
	const u32 n = min(*n_dev, n_max), G = gridDim.x, b = blockIdx.x;
	s_run[threadIdx.x] = base[(size_t)threadIdx.x * G + b];
	u32 lo, hi;
	rs_chunk(n, G, b, lo, hi);
	rs_block_scatter<RS_THREADS, RS_ROUNDS>(keys_in, vals_in, keys_out, vals_out, lo, hi, shift, s_cnt, s_run);
}

size_t ugk_sort_hist_entries(int sms) { return (size_t)256 * sms * RS_BLOCKS_PER_SM; }

int ugk_radix_sort_pairs(cudaStream_t s, u32 *keys[2], u32 *vals[2], const u32 *n_dev, size_t n_max, int bits, SortScratch &sc, int sms, int *launches)
{
	const int G = sms * RS_BLOCKS_PER_SM;
	int cur = 0;
	for (int shift = 0; shift < bits; shift += 8) {
		rs_hist<<<G, RS_THREADS, 0, s>>>(keys[cur], n_dev, (u32)n_max, shift, sc.tile_hist);
		ugk_exclusive_scan(s, sc.tile_hist, sc.tile_base, (size_t)256 * G, sc.scan_state, sc.scan_tiles, sms);
		rs_scatter<<<G, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n_dev, (u32)n_max, shift, sc.tile_base);
		cur ^= 1;
		if (launches) *launches += 3;
	}
	return cur;
}

/* ---------------------------------------------------------------- generic exclusive scan
 * Single pass, decoupled look-back (tiles take tickets in launch order, so a tile only ever waits
 * for tiles that are already running).  state[0] = ticket counter, state[1+tile] = (flag<<32)|value
 * with flag 1 = tile aggregate, 2 = inclusive prefix.  The caller zeroes `state` before the launch. */
#define SC_THREADS 256
#define SC_ITEMS 16
#define SC_TILE (SC_THREADS * SC_ITEMS)

size_t ugk_scan_tiles(size_t n) { return (n + SC_TILE - 1) / SC_TILE + 1; }

__global__ void __launch_bounds__(SC_THREADS) sc_exclusive(const u32 *__restrict__ in, u32 *__restrict__ out, u32 n, u64 *state)
{
	__shared__ u32 s_tile, s_warp[SC_THREADS / 32], s_prefix;
	if (threadIdx.x == 0) s_tile = (u32)atomicAdd((unsigned long long *)&state[0], 1ull);
	__syncthreads();
	const u32 tile = s_tile, base = tile * SC_TILE + threadIdx.x * SC_ITEMS;
	volatile u64 *desc = state + 1;
	u32 v[SC_ITEMS], sum = 0;
#pragma unroll
	for (int k = 0; k < SC_ITEMS; k++) {
		v[k] = (base + k < n) ? in[base + k] : 0u;
		sum += v[k];
	}
	u32 incl = sum;
	for (int o = 1; o < 32; o <<= 1) {
		const u32 x = __shfl_up_sync(UG_FULL_MASK, incl, o);
		if ((int)ug_lane() >= o) incl += x;
	}
	if (ug_lane() == 31) s_warp[threadIdx.x >> 5] = incl;
	__syncthreads();
	/* Sums that can exceed 32 bits (pair counts of a grid whose triangles cover most of a million cells) SATURATE instead of
	 * wrapping: a saturated grand total is larger than any arena, so the frame raises its overflow flag and fails loudly when
	 * the arena cannot grow that far, instead of rendering a truncated list. */
	u64 warp_off = 0, total64 = 0;
#pragma unroll
	for (int w = 0; w < SC_THREADS / 32; w++) {
		const u32 c = s_warp[w];
		if (w < (int)(threadIdx.x >> 5)) warp_off += c;
		total64 += c;
	}
	const u32 total = total64 > 0xffffffffull ? 0xffffffffu : (u32)total64;
	if (threadIdx.x < 32) {
		u64 prefix = 0;
		if (tile == 0) {
			if (threadIdx.x == 0) desc[0] = (2ull << 32) | total;
		} else {
			if (threadIdx.x == 0) desc[tile] = (1ull << 32) | total;
			int look = (int)tile - 1;
			while (true) { /* warp-wide window of 32 predecessors */
				const int idx = look - (int)threadIdx.x;
				u64 d = 0;
				if (idx >= 0) {
					do { d = desc[idx]; } while ((d >> 32) == 0);
				} else {
					d = 2ull << 32; /* before tile 0: inclusive prefix 0 */
				}
				const u32 flag = (u32)(d >> 32), val = (u32)d;
				const u32 incl_mask = __ballot_sync(UG_FULL_MASK, flag == 2);
				if (incl_mask) { /* sum everything up to and including the nearest inclusive prefix */
					const int first = __ffs(incl_mask) - 1;
					u64 part = ((int)threadIdx.x <= first) ? val : 0u;
					for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(UG_FULL_MASK, part, o);
					prefix += part;
					break;
				}
				u64 part = val;
				for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(UG_FULL_MASK, part, o);
				prefix += part;
				look -= 32;
			}
			if (prefix > 0xffffffffull) prefix = 0xffffffffull;
			if (threadIdx.x == 0) desc[tile] = (2ull << 32) | (prefix + total > 0xffffffffull ? 0xffffffffull : prefix + total);
		}
		if (threadIdx.x == 0) s_prefix = (u32)prefix;
	}
	__syncthreads();
	u64 run = (u64)s_prefix + warp_off + incl - sum;
#pragma unroll
	for (int k = 0; k < SC_ITEMS; k++) {
		if (base + k <= n) out[base + k] = run > 0xffffffffull ? 0xffffffffu : (u32)run; /* entry n holds the grand total */
		run += v[k];
	}
}

void ugk_exclusive_scan(cudaStream_t s, const u32 *in, u32 *out, size_t n, u64 *state, size_t state_tiles, int)
{
	const size_t tiles = (n + 1 + SC_TILE - 1) / SC_TILE; /* n+1 outputs */
	cudaMemsetAsync(state, 0, sizeof(u64) * (tiles + 1 < state_tiles + 1 ? tiles + 1 : state_tiles + 1), s);
	sc_exclusive<<<(unsigned)tiles, SC_THREADS, 0, s>>>(in, out, (u32)n, state);
}

__global__ void k_fill_u32(u32 *p, u32 v, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
void ugk_fill_u32(cudaStream_t s, u32 *p, u32 v, size_t n, int sms)
{
	if (!n) return;
	const size_t want = (n + 255) / 256;
	k_fill_u32<<<(unsigned)(want < (size_t)sms * 8 ? want : (size_t)sms * 8), 256, 0, s>>>(p, v, n);
}

// Position finalisation: staging entries [tile | rank-in-tile | pos-in-tile] -> sorted
// absolute positions, without a sort.  Every warp tile reports how many matches it
// produced (tile_count); an exclusive scan over the tiles gives each tile its output
// segment, and every staged match drops into segment_base + rank.
//   kernel A  block-wise exclusive scan of tile_count (in place) + per-block totals
//   kernel B  prefix of the block totals (each CTA, in shared memory) + scatter
#include "scan_common.cuh"

namespace acwm {

constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads) tile_scan_kernel(uint32_t *tile_count, uint64_t n_tiles,
		uint32_t items, unsigned long long *block_sums, const Control *ctl) {
	if (ctl->cursor == 0)
		return; // nothing staged: nothing to place
	__shared__ uint32_t warp_tot[32];
	const uint64_t base = ((uint64_t) blockIdx.x * kScanThreads + threadIdx.x) * items;
	uint32_t sum = 0;
	for (uint32_t i = 0; i < items; i++)
		if (base + i < n_tiles)
			sum += tile_count[base + i];
	const uint32_t incl = warp_incl_scan(sum);
	if (lane_id() == 31)
		warp_tot[threadIdx.x >> 5] = incl;
	__syncthreads();
	if (threadIdx.x < 32) {
		const uint32_t v = warp_tot[threadIdx.x];
		const uint32_t s = warp_incl_scan(v);
		warp_tot[threadIdx.x] = s - v;
		if (threadIdx.x == 31)
			block_sums[blockIdx.x] = s;
	}
	__syncthreads();
	uint32_t run = warp_tot[threadIdx.x >> 5] + incl - sum;
	for (uint32_t i = 0; i < items; i++)
		if (base + i < n_tiles) {
			const uint32_t c = tile_count[base + i];
			tile_count[base + i] = run;
			run += c;
		}
}

__global__ void __launch_bounds__(256) scatter_kernel(const uint64_t *staging, uint64_t cap, const uint32_t *tile_excl,
		const unsigned long long *block_sums, uint32_t n_blocks, uint32_t tiles_per_block, uint32_t tile_syms,
		uint64_t data_lo, uint64_t *positions, Control *ctl) {
	extern __shared__ unsigned long long s_prefix[]; // exclusive prefix of block_sums
	const unsigned long long staged = ctl->cursor < cap ? ctl->cursor : cap;
	if (staged == 0)
		return;
	if (threadIdx.x == 0) {
		unsigned long long run = 0;
		for (uint32_t b = 0; b < n_blocks; b++) {
			s_prefix[b] = run;
			run += block_sums[b];
		}
	}
	__syncthreads();
	const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < staged; i += stride) {
		const uint64_t e = staging[i];
		const uint64_t tile = e >> (kRankBits + kPosBits);
		const uint32_t rank = (uint32_t) (e >> kPosBits) & ((1u << kRankBits) - 1);
		const uint32_t pos = (uint32_t) e & ((1u << kPosBits) - 1);
		const uint64_t at = s_prefix[tile / tiles_per_block] + tile_excl[tile] + rank;
		if (at < cap)
			positions[at] = tile * tile_syms + pos - data_lo;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
		ctl->written = staged;
}

cudaError_t launch_finalize(uint32_t *tile_count, uint64_t n_tiles, unsigned long long *block_sums,
		uint32_t max_blocks, const uint64_t *staging, uint64_t cap, uint32_t tile_syms, uint64_t data_lo,
		uint64_t *positions, Control *ctl, int sm_count, cudaStream_t st) {
	// tiles per scan block: 1024 threads x items, items chosen so that n_blocks <= max_blocks
	uint32_t items = 4;
	while (((n_tiles + (uint64_t) kScanThreads * items - 1) / ((uint64_t) kScanThreads * items)) > max_blocks)
		items *= 2;
	const uint32_t tpb = kScanThreads * items;
	const uint32_t n_blocks = (uint32_t) ((n_tiles + tpb - 1) / tpb);
	tile_scan_kernel<<<n_blocks, kScanThreads, 0, st>>>(tile_count, n_tiles, items, block_sums, ctl);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return e;
	scatter_kernel<<<sm_count * 2, 256, n_blocks * sizeof(unsigned long long), st>>>(staging, cap, tile_count,
			block_sums, n_blocks, tpb, tile_syms, data_lo, positions, ctl);
	return cudaGetLastError();
}

} // namespace acwm

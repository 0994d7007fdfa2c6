/*
 * scan_large.cuh -- N >= 8192 bins (bin_e 13..21): one FFT block per read, too
 * big for one CTA's shared memory (4N bytes = 32 KiB .. 8 MiB).
 *
 * fix_fft (reference src/rtl_power.c:271-327) is a radix-2 DIT on bit-reversed
 * data: stage s pairs positions that differ in bit s.  The stages are run in
 * up to three rounds, each a kernel over 4096-sample tiles that reuses the
 * register-blocked engine of scan_kernels.cuh; between rounds the packed int16
 * data lives in a global scratch (L2 resident: 512 KiB per read at N = 2^17):
 *
 *   round A  stages 0..7     reads the u8 buffer (or a decimated c16 image),
 *                            removes DC, applies the window, and performs the
 *                            bit-reversal as a tile transpose: a tile is 16
 *                            consecutive input samples x 256 strided rows, so
 *                            both the gather and the scatter move >= 32-byte runs
 *   round B  stages 8..15    in place on the scratch, tile = consecutive
 *                            positions x 2^LB strided rows
 *   round C  stages 16..L-1  (N > 65536) register-only, 2^(L-16) strided points
 *                            per thread, then |X|^2 accumulation in natural bin
 *                            order (coalesced 64-bit atomics)
 * The last round accumulates |X|^2 / peak hold (rtl_power.c:708-716).
 * Rounding, halving, twiddles and int16 wrap are those of butterfly().
 */
#pragma once
#include "scan_kernels.cuh"

namespace rscan {

struct LargeParams {
	const uint8_t *base;        /* u8 reads or decimated c16 images */
	const long long *read_off;  /* byte offset per entry, NULL = regular */
	long long regular_stride;
	int entry_base;             /* first entry of this chunk */
	const int *hop_of;          /* hop per entry (absolute entry index) */
	c16 *scratch;               /* [chunk][N] */
	const long long *dc_sums;   /* [chunk][2] */
	long long *avg;
	long long *samples;         /* [tune_count] */
	int samples_per_read;
	const int2 *tw;             /* [N/2] */
	const int2 *twb;            /* round-B re-ordered copy, see TwLargeB */
	const uint16_t *win;        /* [N] */
	int L;
	int n_entries;              /* entries in this chunk */
	PassTw tw0;
};

SCAN_DEV long long large_entry_offset(const LargeParams &prm, int e)
{
	return prm.read_off ? prm.read_off[e] : (long long)(e - prm.entry_base) * prm.regular_stride;
}

/* per-read byte sums of I and Q for the u8 input (remove_dc, rtl_power.c:586-588) */
struct DcSumU8Params {
	const uint8_t *base;
	const long long *read_off;
	int entry_base;
	int buf_len;
	long long *sums; /* [chunk][2], zeroed by the host */
};

__global__ void __launch_bounds__(256)
dc_sums_u8_kernel(const SCAN_GRID_CONSTANT DcSumU8Params prm)
{
	const int rel = blockIdx.y;
	const uint8_t *src = prm.base + prm.read_off[prm.entry_base + rel];
	unsigned sI = 0, sQ = 0; /* <= 2^21 bytes of 255 per component: fits */
	const int stride = gridDim.x * blockDim.x * 16;
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 16; i < prm.buf_len; i += 8 * stride) {
		uint4 q[8]; /* 128 bytes in flight per thread: one resident wave of CTAs streams at HBM speed */
#pragma unroll
		for (int j = 0; j < 8; ++j)
			q[j] = (i + j * stride < prm.buf_len) ? __ldg((const uint4 *)(src + i + j * stride)) : uint4{ 0, 0, 0, 0 };
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			sI = __dp4a(q[j].x, 0x00010001u, sI);
			sQ = __dp4a(q[j].x, 0x01000100u, sQ);
			sI = __dp4a(q[j].y, 0x00010001u, sI);
			sQ = __dp4a(q[j].y, 0x01000100u, sQ);
			sI = __dp4a(q[j].z, 0x00010001u, sI);
			sQ = __dp4a(q[j].z, 0x01000100u, sQ);
			sI = __dp4a(q[j].w, 0x00010001u, sI);
			sQ = __dp4a(q[j].w, 0x01000100u, sQ);
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		sI += __shfl_xor_sync(0xffffffffu, sI, o);
		sQ += __shfl_xor_sync(0xffffffffu, sQ, o);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd((unsigned long long *)(prm.sums + 2 * rel), (unsigned long long)sI);
		atomicAdd((unsigned long long *)(prm.sums + 2 * rel + 1), (unsigned long long)sQ);
	}
}

/* ---- round A ----------------------------------------------------------- */

struct TwLargeA {
	static constexpr bool kTrivial = true;
	const int2 *twc;     /* compact: stage s (4..7), group m at twc[(1<<s)-16+m] */
	const PassTw *tw0;
	template <int K>
	SCAN_DEV int2 get(int s, int pa) const
	{
		const int m = pa & ((1 << s) - 1);
		if constexpr (K == 0)
			return tw0->w[(1 << s) - 1 + m];
		else
			return twc[(1 << s) - 16 + m];
	}
};

constexpr int kLargeSmemA = kXchWords * 4 + 240 * 8 + 16;

template <bool IN16>
__global__ void __launch_bounds__(kThreads, 2)
large_round_a_kernel(const SCAN_GRID_CONSTANT LargeParams prm)
{
	SCAN_DYN_SMEM(smem);
	c16 *stage = (c16 *)smem;
	int2 *twc = (int2 *)(smem + kXchWords * 4);
	int *dck = (int *)(smem + kXchWords * 4 + 240 * 8);
	const int t = threadIdx.x, L = prm.L;
	const int tile = blockIdx.x, rel = blockIdx.y, e = prm.entry_base + rel;
	const long long N = 1ll << L;
	const uint8_t *src = prm.base + large_entry_offset(prm, e);

	if (t < 240) {
		/* entry t of the compact table: stage s = 4 + floor(log2(t/16 + 1)) */
		int s = 4, off = 0;
		while (t >= off + (1 << s)) {
			off += 1 << s;
			s++;
		}
		twc[t] = prm.tw[(size_t)(t - off) << (L - 1 - s)];
	}
	if (t == 0 && tile == 0)
		atomicAdd((unsigned long long *)(prm.samples + prm.hop_of[e]), (unsigned long long)prm.samples_per_read);
	if (t < 2) {
		long long s = prm.dc_sums[2 * rel + t];
		if constexpr (!IN16)
			dck[t] = 127 + dc_average(s - 127ll * N, (int)(2 * N) - t);
		else
			dck[t] = dc_average(s, (int)(2 * N) - t);
	}

	__syncthreads();
	const int kI = dck[0], kQ = dck[1];

	/* gather: thread t owns row i = t = 16 consecutive input samples and their 16 window
	 * coefficients (two 16-byte loads each); convert, remove DC, window, and park the c16
	 * results column-major so that the engine finds column c, row i at stage[c*256 + i] */
	{
		const long long n0 = ((long long)t << (L - 8)) + 16 * tile;
		const uint4 wa = __ldg((const uint4 *)(prm.win + n0));
		const uint4 wb = __ldg((const uint4 *)(prm.win + n0 + 8));
		const unsigned ww[8] = { wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w };
		if constexpr (!IN16) {
			const uint4 a = __ldg((const uint4 *)(src + 2 * n0));
			const uint4 b = __ldg((const uint4 *)(src + 2 * n0 + 16));
			const unsigned w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
			for (int c = 0; c < 16; ++c) {
				const unsigned raw = (w[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
				const int wv = (int)((ww[c >> 1] >> (16 * (c & 1))) & 0xFFFFu);
				stage[xch_idx(c * 256 + t)] = c16_pack(((int)(raw & 0xFFu) - kI) * wv, ((int)(raw >> 8) - kQ) * wv);
			}
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const uint4 a = __ldg((const uint4 *)(src + 4 * n0 + 16 * k));
				const unsigned raw[4] = { a.x, a.y, a.z, a.w };
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const int c = 4 * k + j;
					const int wv = (int)((ww[c >> 1] >> (16 * (c & 1))) & 0xFFFFu);
					stage[xch_idx(c * 256 + t)] = c16_pack((c16_re(raw[j]) - kI) * wv, (c16_im(raw[j]) - kQ) * wv);
				}
			}
		}
	}
	__syncthreads();

	/* position 16t + r is column c = t >> 4, row bitrev8(q) */
	X2 x[kPts];
	const int c = t >> 4;
#pragma unroll
	for (int r = 0; r < kPts; ++r) {
		const int i = (brev4(r) << 4) | brev_bits((unsigned)(t & 15), 4);
		x[r] = x_unpack(stage[xch_idx(c * 256 + i)]);
	}

	TwLargeA tw;
	tw.twc = twc;
	tw.tw0 = &prm.tw0;
	engine_fft<8>(x, stage, t, tw);

	/* scatter: column n_low lands on positions (bitrev(n_low) << 8) | q */
	c16 *dst = prm.scratch + (long long)rel * N;
#pragma unroll
	for (int r = 0; r < kPts; ++r) {
		const int p = last_pos<8>(t, r);
		const int cc = p >> 8, q = p & 255;
		const long long P = ((long long)brev_bits((unsigned)(16 * tile + cc), L - 8) << 8) | q;
		dst[P] = x_pack(x[r]);
	}
}

/* ---- round B ----------------------------------------------------------- */

/*
 * Round-B twiddles.  Stage 8+se pairs positions whose low 8+se bits m = (ilow << 8) | plow
 * select tw[m << (L-9-se)].  For se >= 4 the lanes of a warp differ in ilow, i.e. their
 * entries are 2^(L-1-se) apart in tw[]: one 32-byte sector per lane.  The host therefore
 * also provides the same values re-ordered as twb[se][plow][ilow] (ilow fastest), so that a
 * warp reads one or two contiguous 128-byte runs.
 */
SCAN_DEV constexpr long long twb_offset(int se) { return 256ll * ((1 << se) - 16); } /* se >= 4 */

template <int LB>
struct TwLargeB {
	static constexpr bool kTrivial = false;
	const int2 *tw;
	const int2 *twb;
	int plow0, L;
	template <int K>
	SCAN_DEV int2 get(int se, int pa) const
	{
		const int col = pa >> LB, i = pa & ((1 << LB) - 1);
		const int ilow = i & ((1 << se) - 1);
		if (K >= 1)
			return __ldg(twb + twb_offset(se) + ((long long)(plow0 + col) << se) + ilow);
		const long long m = ((long long)ilow << 8) | (plow0 + col);
		return __ldg(tw + (m << (L - 9 - se)));
	}
};

SCAN_DEV void accumulate_bin(long long *dst, c16 x, bool peak)
{
	const int re = c16_re(x), im = c16_im(x);
	const unsigned pw = (unsigned)(re * re) + (unsigned)(im * im);
	if (peak)
		atomicMax(dst, (long long)pw);
	else
		atomicAdd((unsigned long long *)dst, (unsigned long long)pw);
}

/* Tile transposes of round B: columns are 2^LB words apart, so on top of the engine's
 * p>>4 padding the column index is spread over the banks with 2 * (p >> 8). */
SCAN_DEV int tile_idx(int p) { return p + (p >> 4) + 2 * (p >> 8); }
constexpr int kLargeSmemB = (kWS + kWS / 16 + 2 * (kWS / 256) + 16) * 4;

template <int LB, bool LAST, bool PEAK>
__global__ void __launch_bounds__(kThreads, 2) /* (3 CTAs per SM at 80 registers: measured 103 vs 105 G samples/s) */
large_round_b_kernel(const SCAN_GRID_CONSTANT LargeParams prm)
{
	SCAN_DYN_SMEM(smem);
	c16 *stage = (c16 *)smem;
	constexpr int ncols = kWS >> LB;
	constexpr int tiles_per_u = 256 / ncols;
	const int t = threadIdx.x, L = prm.L;
	const int rel = blockIdx.y;
	const long long N = 1ll << L;
	const int U = blockIdx.x / tiles_per_u, plow0 = (blockIdx.x % tiles_per_u) * ncols;
	c16 *data = prm.scratch + (long long)rel * N + ((long long)U << (8 + LB)) + plow0;

#pragma unroll
	for (int k = 0; k < kPts; ++k) {
		const int eidx = t + kThreads * k;
		const int col = eidx % ncols, i = eidx / ncols;
		stage[tile_idx((col << LB) | i)] = data[((long long)i << 8) + col];
	}
	__syncthreads();
	X2 x[kPts];
#pragma unroll
	for (int r = 0; r < kPts; ++r)
		x[r] = x_unpack(stage[tile_idx(pos<0>(t, r))]);

	TwLargeB<LB> tw;
	tw.tw = prm.tw;
	tw.twb = prm.twb;
	tw.plow0 = plow0;
	tw.L = L;
	engine_fft<LB>(x, stage, t, tw);

	__syncthreads();
#pragma unroll
	for (int r = 0; r < kPts; ++r)
		stage[tile_idx(last_pos<LB>(t, r))] = x_pack(x[r]);
	__syncthreads();
	long long *out = nullptr;
	if constexpr (LAST)
		out = prm.avg + ((long long)prm.hop_of[prm.entry_base + rel] << L) + plow0;
#pragma unroll
	for (int k = 0; k < kPts; ++k) {
		const int eidx = t + kThreads * k;
		const int col = eidx % ncols, i = eidx / ncols;
		const c16 x = stage[tile_idx((col << LB) | i)];
		if constexpr (LAST)
			accumulate_bin(out + ((long long)i << 8) + col, x, PEAK);
		else
			data[((long long)i << 8) + col] = x;
	}
}

/* ---- round C ----------------------------------------------------------- */

constexpr int kRoundCReads = 16; /* reads one CTA of round C walks through */

template <int LC, bool PEAK>
__global__ void __launch_bounds__(kThreads)
large_round_c_kernel(const SCAN_GRID_CONSTANT LargeParams prm)
{
	constexpr int R = 1 << LC;
	constexpr bool kHoist = LC <= 3; /* twiddles depend on the position only: keep them in registers */
	const int L = prm.L;
	const int plow = blockIdx.x * kThreads + threadIdx.x; /* 0 .. 65535 */
	const long long N = 1ll << L;
	int2 wreg[kHoist ? R - 1 : 1];
	if constexpr (kHoist) {
#pragma unroll
		for (int se = 0; se < LC; ++se)
#pragma unroll
			for (int g = 0; g < (1 << se); ++g) {
				const long long m = ((long long)g << 16) | plow;
				wreg[(1 << se) - 1 + g] = __ldg(prm.tw + (m << (L - 17 - se)));
			}
	}
	unsigned long long acc[R];
#pragma unroll
	for (int r = 0; r < R; ++r)
		acc[r] = 0ull;
	int cur_hop = -1;
	const int rel0 = blockIdx.y * kRoundCReads;
	const int rel1 = (rel0 + kRoundCReads < prm.n_entries) ? rel0 + kRoundCReads : prm.n_entries;
	for (int rel = rel0; rel <= rel1; ++rel) {
		const int hop = (rel < rel1) ? prm.hop_of[prm.entry_base + rel] : -1;
		if (hop != cur_hop) {
			if (cur_hop >= 0) {
				long long *out = prm.avg + ((long long)cur_hop << L) + plow;
#pragma unroll
				for (int r = 0; r < R; ++r) {
					if (PEAK)
						atomicMax(out + ((long long)r << 16), (long long)acc[r]);
					else
						atomicAdd((unsigned long long *)(out + ((long long)r << 16)), acc[r]);
					acc[r] = 0ull;
				}
			}
			cur_hop = hop;
		}
		if (rel == rel1)
			break;
		const c16 *data = prm.scratch + (long long)rel * N + plow;
		c16 v[R];
#pragma unroll
		for (int r = 0; r < R; ++r)
			v[r] = data[(long long)r << 16];
#pragma unroll
		for (int se = 0; se < LC; ++se) {
#pragma unroll
			for (int r = 0; r < R; ++r) {
				if ((r & (1 << se)) == 0) {
					const int g = r & ((1 << se) - 1);
					int2 w;
					if constexpr (kHoist) {
						w = wreg[(1 << se) - 1 + g];
					} else {
						const long long m = ((long long)g << 16) | plow;
						w = __ldg(prm.tw + (m << (L - 17 - se)));
					}
					butterfly(v[r], v[r | (1 << se)], w.x, w.y);
				}
			}
		}
#pragma unroll
		for (int r = 0; r < R; ++r) {
			const int re = c16_re(v[r]), im = c16_im(v[r]);
			accumulate_power<PEAK>(acc[r], re, im);
		}
	}
}


/* ======================================================================== *
 *  Two-round path for 2^13 .. 2^17 bins (round 2)                           *
 * ======================================================================== */

/*
 * The three-round path above spends ~215 instructions per sample at 2^17 bins where the stages themselves need
 * 17 x 8: every round pays a load / transpose / store of its own, and rounds A and B are one-tile CTAs with
 * nothing in flight while they transform.  Here the stages are cut 12 + (L - 12):
 *
 *   permute   (memory bound)  u8 -> packed int16 (minus 127, rtl_power.c:666-668) written at its bit-reversed
 *                             position (fix_fft's permutation, :282-297) as a tile transpose with 32/64-byte
 *                             runs on both sides; also the byte sums remove_dc needs (:586-588) -- no separate
 *                             pass over the input
 *   mid       stages 0..11    persistent CTAs, equal runs of 4096-position tiles (16 KiB, contiguous), the next
 *                             tile in flight (cp.async) while this one is transformed with the register-blocked
 *                             engine of scan_kernels.cuh; DC + window applied on the way in (window table stored
 *                             in position order), result written back in place
 *   top       stages 12..L-1  tile = 2^(L-12) rows 4096 apart x (4096 >> (L-12)) consecutive columns; a CTA keeps
 *                             its tile's twiddles in shared memory and walks several reads, |X|^2 accumulated in
 *                             registers (rtl_power.c:708-716), one coalesced 64-bit RED per bin and hop
 * Rounding, halving, twiddles and int16 wrap are those of butterfly() at every stage.
 */
struct Large2Params {
	const uint8_t *base;        /* u8 reads or decimated c16 images */
	const long long *read_off;  /* byte offset per entry, NULL = regular */
	long long regular_stride;
	int entry_base;
	const int *hop_of;
	c16 *scratch;               /* [chunk][N], position order */
	long long *sums;            /* [chunk][2] byte sums (u8) / sample sums (c16), zeroed by the host for u8 */
	long long *avg;
	long long *samples;
	int samples_per_read;
	const uint16_t *wperm;      /* [N] window coefficients in POSITION order: wperm[p] = win[bitrev_L(p)] */
	const int2 *twc12;          /* stages 4..11, group m of stage s at (1<<s)-16+m = tw[m << (L-1-s)] */
	const int2 *twt;            /* top stages: stage 12+se, [ilow < 2^se][plow < 4096] at 4096*((1<<se)-1) + (ilow<<12) + plow */
	int L;
	int n_entries;
	int top_reads;              /* reads one CTA of the top kernel walks */
	int in16;                   /* input is a decimated c16 image (sums come from the decimators) */
	PassTw tw0;
};

SCAN_DEV long long large2_entry_offset(const Large2Params &prm, int e)
{
	return prm.read_off ? prm.read_off[e] : (long long)(e - prm.entry_base) * prm.regular_stride;
}

/* ---- permute ----------------------------------------------------------- */

constexpr int kLarge2SmemPermute = kXchWords * 4;

template <bool IN16>
__global__ void __launch_bounds__(kThreads)
large2_permute_kernel(const SCAN_GRID_CONSTANT Large2Params prm)
{
	SCAN_DYN_SMEM(smem);
	c16 *stage = (c16 *)smem;
	const int t = threadIdx.x, L = prm.L;
	const int tile = blockIdx.x, rel = blockIdx.y, e = prm.entry_base + rel;
	const long long N = 1ll << L;
	const uint8_t *src = prm.base + large2_entry_offset(prm, e);

	if (t == 0 && tile == 0)
		atomicAdd((unsigned long long *)(prm.samples + prm.hop_of[e]), (unsigned long long)prm.samples_per_read);
	/* gather: thread t owns row t = 16 consecutive input samples; park them column-major */
	{
		const long long n0 = ((long long)t << (L - 8)) + 16 * tile;
		if constexpr (!IN16) {
			const uint4 a = __ldg((const uint4 *)(src + 2 * n0));
			const uint4 b = __ldg((const uint4 *)(src + 2 * n0 + 16));
			const unsigned w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
			unsigned sI = 0, sQ = 0;
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				sI = __dp4a(w[k], 0x00010001u, sI);
				sQ = __dp4a(w[k], 0x01000100u, sQ);
			}
#pragma unroll
			for (int c = 0; c < 16; ++c) {
				const unsigned raw = (w[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
				stage[xch_idx(c * 256 + t)] = c16_pack((int)(raw & 0xFFu) - 127, (int)(raw >> 8) - 127);
			}
			/* byte sums of the whole read (remove_dc, rtl_power.c:586-588): one 64-bit atomic pair per warp */
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				sI += __shfl_xor_sync(0xffffffffu, sI, o);
				sQ += __shfl_xor_sync(0xffffffffu, sQ, o);
			}
			if ((t & 31) == 0) {
				atomicAdd((unsigned long long *)(prm.sums + 2 * rel), (unsigned long long)sI);
				atomicAdd((unsigned long long *)(prm.sums + 2 * rel + 1), (unsigned long long)sQ);
			}
		} else {
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const uint4 a = __ldg((const uint4 *)(src + 4 * n0 + 16 * k));
				stage[xch_idx((4 * k + 0) * 256 + t)] = a.x;
				stage[xch_idx((4 * k + 1) * 256 + t)] = a.y;
				stage[xch_idx((4 * k + 2) * 256 + t)] = a.z;
				stage[xch_idx((4 * k + 3) * 256 + t)] = a.w;
			}
		}
	}
	__syncthreads();
	/* scatter: sample (row i, column cc) of this tile has index n = (i << (L-8)) + 16*tile + cc, its position is
	 * bitrev_L(n) = (bitrev_{L-8}(16*tile + cc) << 8) | bitrev8(i).  Register r of thread t writes column
	 * cc = t >> 4, position low byte q = (r << 4) | (t & 15): 16 consecutive words per half warp */
	c16 *dst = prm.scratch + (long long)rel * N;
	const int cc = t >> 4;
	const long long phi = (long long)brev_bits((unsigned)(16 * tile + cc), L - 8) << 8;
#pragma unroll
	for (int r = 0; r < kPts; ++r) {
		const int q = (r << 4) | (t & 15);
		const int i = brev_bits((unsigned)q, 8);
		dst[phi | q] = stage[xch_idx(cc * 256 + i)];
	}
}

/* ---- mid: stages 0..11 on contiguous 4096-position tiles ------------------ */

struct Large2MidSmem {
	static constexpr int slot_bytes = kWS * 4 + kWS * 2;          /* tile data + its window slice */
	static constexpr int off_slot = 0;                            /* two slots; the current one doubles as a transpose buffer */
	static constexpr int off_xch = 2 * slot_bytes;                /* the other transpose buffer */
	static constexpr int off_tw = off_xch + kXchWords * 4;
	static constexpr int off_dck = off_tw + (kWS - 16) * 8;       /* [2 items][2] DC constants */
	static constexpr int bytes = off_dck + 32;
};
static_assert(Large2MidSmem::slot_bytes >= kXchWords * 4, "a slot must hold a padded transpose buffer");

struct TwMid {
	static constexpr bool kTrivial = true;
	const int2 *tws;
	const PassTw *tw0;
	template <int K>
	SCAN_DEV int2 get(int s, int pa) const
	{
		const int m = pa & ((1 << s) - 1);
		if constexpr (K == 0)
			return tw0->w[(1 << s) - 1 + m];
		else
			return tws[(1 << s) - 16 + m];
	}
};

__global__ void __launch_bounds__(kThreads, 2)
large2_mid_kernel(const SCAN_GRID_CONSTANT Large2Params prm)
{
	SCAN_DYN_SMEM(smem);
	typedef Large2MidSmem SM;
	c16 *xchb = (c16 *)(smem + SM::off_xch);
	int2 *tws = (int2 *)(smem + SM::off_tw);
	int *dck = (int *)(smem + SM::off_dck);
	const int t = threadIdx.x, L = prm.L;
	const long long N = 1ll << L;
	const int tiles = (int)(N / kWS);
	const long long total = (long long)prm.n_entries * tiles;
	const long long w_lo = total * blockIdx.x / gridDim.x, w_hi = total * (blockIdx.x + 1) / gridDim.x;
	if (w_lo >= w_hi)
		return;

	for (int i = t; i < (kWS - 16) / 2; i += kThreads)
		cp_async16((uint8_t *)tws + 16 * i, (const uint8_t *)prm.twc12 + 16 * i);
	TwMid tw;
	tw.tws = tws;
	tw.tw0 = &prm.tw0;

	auto prefetch = [&](long long w, int slot) {
		const int rel = (int)(w / tiles), tile = (int)(w - (long long)rel * tiles);
		const uint8_t *d = (const uint8_t *)(prm.scratch + (long long)rel * N + (long long)tile * kWS);
		const uint8_t *wv = (const uint8_t *)(prm.wperm + (long long)tile * kWS);
		uint8_t *dst = smem + SM::off_slot + slot * SM::slot_bytes;
#pragma unroll
		for (int i = 0; i < kWS * 4 / (kThreads * 16); ++i)
			cp_async16(dst + (i * kThreads + t) * 16, d + (i * kThreads + t) * 16);
#pragma unroll
		for (int i = 0; i < kWS * 2 / (kThreads * 16); ++i)
			cp_async16(dst + kWS * 4 + (i * kThreads + t) * 16, wv + (i * kThreads + t) * 16);
		cp_async_commit();
	};
	/* the int16 averages remove_dc subtracts (rtl_power.c:581-596: divisors 2N and 2N - 1, truncating), as seen by
	 * the stored (b - 127) / decimated values; computed by two threads one item ahead */
	auto dc_of = [&](long long w, int which) {
		const int rel = (int)(w / tiles);
		if (t < 2) {
			const long long s = prm.sums[2 * rel + t];
			dck[which * 2 + t] = prm.in16 ? dc_average(s, (int)(2 * N) - t) : dc_average(s - 127ll * N, (int)(2 * N) - t);
		}
	};

	prefetch(w_lo, 0);
	dc_of(w_lo, 0);
	for (long long w = w_lo; w < w_hi; ++w) {
		const int u = (int)(w - w_lo), slot = u & 1;
		cp_async_wait_all();
		__syncthreads(); /* slot landed (and its DC constants); the other slot is no longer read as a transpose buffer */
		if (w + 1 < w_hi) {
			prefetch(w + 1, slot ^ 1);
			dc_of(w + 1, slot ^ 1);
		}
		const int rel = (int)(w / tiles), tile = (int)(w - (long long)rel * tiles);
		c16 *cur = (c16 *)(smem + SM::off_slot + slot * SM::slot_bytes);
		const uint16_t *wsl = (const uint16_t *)(cur + kWS);
		const int kI = dck[slot * 2], kQ = dck[slot * 2 + 1];

		/* thread t takes positions 16t .. 16t+15: 64 contiguous bytes of samples, 32 of coefficients */
		X2 x[kPts];
		{
			const uint4 *dp = (const uint4 *)(cur + 16 * t);
			const uint4 *wp = (const uint4 *)(wsl + 16 * t);
			const uint4 wa = wp[0], wb = wp[1];
			const unsigned ww[8] = { wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w };
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const uint4 q = dp[k];
				const unsigned v[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const int r = 4 * k + j;
					const unsigned wx = (r & 1) ? (ww[r >> 1] & 0xFFFF0000u) : (ww[r >> 1] << 16);
					x[r].re = (int)((unsigned)(c16_re(v[j]) - kI) * wx);
					x[r].im = (int)((unsigned)(c16_im(v[j]) - kQ) * wx);
				}
			}
		}
		c16 pk[kPts];
		run_pass<0, 12>(x, t, tw);
		exchange_pk<0, 1, BlockBar>(x, pk, xchb, t, t, BlockBar());
		run_pass<1, 12>(x, t, tw, pk);
		exchange_pk<1, 2, BlockBar>(x, pk, cur, t, t, BlockBar()); /* the consumed slot is the second transpose buffer */
		run_pass<2, 12>(x, t, tw, pk);

		c16 *out = prm.scratch + (long long)rel * N + (long long)tile * kWS;
#pragma unroll
		for (int r = 0; r < kPts; ++r)
			out[last_pos<12>(t, r)] = x_pack(x[r]);
	}
}

/* ---- top: stages 12..L-1 + |X|^2 ---------------------------------------- */

template <int LT>
struct Large2TopSmem {
	static constexpr int off_stage = 0;                             /* two 16 KiB tiles */
	static constexpr int off_xch = 2 * kWS * 4;                     /* LT = 5 only: one transpose buffer, and (same
	                                                                 * memory, 4096 x u64) the flush's transpose */
	static constexpr int off_tw = off_xch + (LT > 4 ? kWS * 8 : 0);
	static constexpr int tw_entries = ((1 << LT) - 1) * (kWS >> LT); /* stage se: [ilow][col] at ((1<<se)-1)*ncols */
	static constexpr int bytes = off_tw + tw_entries * 8 + 16;
};

template <int LT>
struct TwTop {
	static constexpr bool kTrivial = false;
	const int2 *tws;
	/* engine position pa = (col << LT) | i; stage se pairs i and i + 2^se, twiddle group m = (i mod 2^se, plow) */
	template <int K>
	SCAN_DEV int2 get(int se, int pa) const
	{
		constexpr int ncols = kWS >> LT;
		const int col = pa >> LT, ilow = pa & ((1 << se) - 1);
		return tws[((1 << se) - 1) * ncols + ilow * ncols + col];
	}
};

template <int LT, bool PEAK>
__global__ void __launch_bounds__(kThreads, 2)
large2_top_kernel(const SCAN_GRID_CONSTANT Large2Params prm)
{
	SCAN_DYN_SMEM(smem);
	typedef Large2TopSmem<LT> SM;
	constexpr int ncols = kWS >> LT, rows = 1 << LT;
	c16 *xch = (c16 *)(smem + SM::off_xch);
	int2 *tws = (int2 *)(smem + SM::off_tw);
	const int t = threadIdx.x, L = prm.L;
	const long long N = 1ll << L;
	const int plow0 = blockIdx.x * ncols;
	const int rel0 = blockIdx.y * prm.top_reads;
	const int rel1 = (rel0 + prm.top_reads < prm.n_entries) ? rel0 + prm.top_reads : prm.n_entries;
	if (rel0 >= rel1)
		return;

	/* this tile's twiddles, once per CTA: stage 12+se, (ilow, col) <- twt[4096*((1<<se)-1) + (ilow << 12) + plow0 + col] */
	for (int idx = t; idx < SM::tw_entries; idx += kThreads) {
		int se = 0, off = 0;
		while (idx >= off + (ncols << se)) {
			off += ncols << se;
			se++;
		}
		const int ilow = (idx - off) / ncols, col = (idx - off) % ncols;
		tws[idx] = __ldg(prm.twt + 4096ll * ((1 << se) - 1) + ((long long)ilow << 12) + plow0 + col);
	}
	TwTop<LT> tw;
	tw.tws = tws;

	/* tile of read `rel`: row i (0 .. 2^LT-1) = ncols consecutive words at scratch[rel][(i << 12) + plow0] */
	auto prefetch = [&](int rel, int slot) {
		const uint8_t *d = (const uint8_t *)(prm.scratch + (long long)rel * N + plow0);
		uint8_t *dst = smem + SM::off_stage + slot * kWS * 4;
		constexpr int chunks_per_row = ncols * 4 / 16;
#pragma unroll
		for (int k = 0; k < kWS * 4 / (kThreads * 16); ++k) {
			const int c = k * kThreads + t;
			const int i = c / chunks_per_row, j = c - i * chunks_per_row;
			cp_async16(dst + ((long long)i * ncols * 4) + j * 16, d + ((long long)i << 14) + j * 16);
		}
		cp_async_commit();
	};

	unsigned long long acc[kPts];
#pragma unroll
	for (int r = 0; r < kPts; ++r)
		acc[r] = 0ull;
	int cur_hop = prm.hop_of[prm.entry_base + rel0];
	bool waited = false;
	prefetch(rel0, 0);
	for (int rel = rel0; rel < rel1; ++rel) {
		const int slot = (rel - rel0) & 1;
		cp_async_wait_all();
		__syncthreads();
		if (rel + 1 < rel1)
			prefetch(rel + 1, slot ^ 1);
		const c16 *cur = (const c16 *)(smem + SM::off_stage + slot * kWS * 4);
		/* engine position 16t + r = (col << LT) | i  ->  staged word i * ncols + col */
		X2 x[kPts];
#pragma unroll
		for (int r = 0; r < kPts; ++r) {
			const int p = 16 * t + r, col = p >> LT, i = p & (rows - 1);
			x[r] = x_unpack(cur[i * ncols + col]);
		}
		run_pass<0, LT>(x, t, tw);
		if constexpr (LT > 4) {
			exchange<0, 1, true>(x, xch, t, t);
			run_pass<1, LT>(x, t, tw);
		}
#pragma unroll
		for (int r = 0; r < kPts; ++r)
			accumulate_power<PEAK>(acc[r], x[r].re >> 16, x[r].im >> 16);

		const int next_hop = (rel + 1 < rel1) ? prm.hop_of[prm.entry_base + rel + 1] : -1;
		if (next_hop != cur_hop) {
			if (!waited)
				pdl_wait();
			waited = true;
			long long *out = prm.avg + ((long long)cur_hop << L) + plow0;
			if constexpr (LT <= 4) {
				/* one pass: register r of thread t holds (col, i) = (p >> LT, p mod 2^LT), p = 16t + r: for a
				 * fixed r the lanes' bins are at most 2^(4-LT) words apart -> coalesced as they are */
#pragma unroll
				for (int r = 0; r < kPts; ++r) {
					const int p = last_pos<LT>(t, r), col = p >> LT, i = p & (rows - 1);
					long long *dst = out + ((long long)i << 12) + col;
					if (PEAK)
						atomicMax(dst, (long long)acc[r]);
					else
						atomicAdd((unsigned long long *)dst, acc[r]);
					acc[r] = 0ull;
				}
			} else {
				/* after the second pass the lanes of a warp hold rows 4096 bins apart: transpose the sums through
				 * shared memory (the transpose buffer, idle here) so that a warp's atomics cover 32 consecutive bins */
				unsigned long long *fb = (unsigned long long *)(smem + SM::off_xch);
				__syncthreads();
#pragma unroll
				for (int r = 0; r < kPts; ++r) {
					const int p = last_pos<LT>(t, r), col = p >> LT, i = p & (rows - 1);
					fb[i * ncols + col] = acc[r];
					acc[r] = 0ull;
				}
				__syncthreads();
#pragma unroll
				for (int k = 0; k < kPts; ++k) {
					const int idx = t + kThreads * k, i = idx / ncols, col = idx - i * ncols;
					long long *dst = out + ((long long)i << 12) + col;
					if (PEAK)
						atomicMax(dst, (long long)fb[idx]);
					else
						atomicAdd((unsigned long long *)dst, fb[idx]);
				}
				/* (the next transpose into this buffer starts with a barrier: exchange<0, 1, true>) */
			}
			cur_hop = next_hop;
		}
	}
}

} // namespace rscan

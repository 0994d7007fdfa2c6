static int g_emu_large3 = 0;
void emu_set_large3(int v) { g_emu_large3 = v; }

/* two-round path (bin_e 13..17): permute -> stages 0..11 -> stages 12..L-1 */
static void emu_large2(int L, int peak, int in16, const uint8_t *reads, int n_reads, const int *hop_of,
		       const int2 *tw, const uint16_t *win, const long long *sums_in, long long *avg, int mid_grid, int top_reads)
{
	const size_t N = (size_t)1 << L;
	const int tiles = (int)(N / kWS), lt = L - 12;
	std::vector<long long> offs(n_reads);
	for (int i = 0; i < n_reads; i++)
		offs[i] = (long long)i * (long long)(in16 ? 4 * N : 2 * N);
	std::vector<c16> scratch((size_t)n_reads * N, 0xDEADBEEFu);
	std::vector<long long> sums((size_t)n_reads * 2, 0);
	if (in16)
		memcpy(sums.data(), sums_in, sums.size() * 8);
	std::vector<int2> twc((size_t)kWS - 16);
	for (int st = 4; st < 12; st++)
		for (int m = 0; m < (1 << st); m++)
			twc[(size_t)(1 << st) - 16 + m] = tw[(size_t)m << (L - 1 - st)];
	std::vector<int2> twt((size_t)4096 * ((1 << lt) - 1));
	for (int se = 0; se < lt; se++)
		for (int ilow = 0; ilow < (1 << se); ilow++)
			for (int plow = 0; plow < 4096; plow++)
				twt[(size_t)4096 * ((1 << se) - 1) + ((size_t)ilow << 12) + plow] = tw[((size_t)(ilow << 12) | plow) << (L - 13 - se)];
	std::vector<uint16_t> wperm(N);
	for (size_t pidx = 0; pidx < N; pidx++) {
		size_t n = 0;
		for (int b = 0; b < L; b++)
			n |= ((pidx >> b) & 1u) << (L - 1 - b);
		wperm[pidx] = win[n];
	}
	std::vector<long long> smp(4096, 0);
	Large2Params p;
	memset(&p, 0, sizeof(p));
	p.base = reads;
	p.read_off = offs.data();
	p.entry_base = 0;
	p.hop_of = hop_of;
	p.scratch = scratch.data();
	p.sums = sums.data();
	p.avg = avg;
	p.samples = smp.data();
	p.samples_per_read = 1;
	p.wperm = wperm.data();
	p.twc12 = twc.data();
	p.twt = twt.data();
	p.L = L;
	p.n_entries = n_reads;
	p.top_reads = top_reads;
	p.in16 = in16;
	fill_tw0(p.tw0, tw, L);
	dim3 tiles_grid((unsigned)tiles, n_reads);
	if (in16)
		cuda_emu::launch(tiles_grid, dim3(kThreads), kLarge2SmemPermute, [&]() { large2_permute_kernel<true>(p); });
	else
		cuda_emu::launch(tiles_grid, dim3(kThreads), kLarge2SmemPermute, [&]() { large2_permute_kernel<false>(p); });
	cuda_emu::launch(dim3(mid_grid), dim3(kThreads), Large2MidSmem::bytes, [&]() { large2_mid_kernel(p); });
	dim3 g((unsigned)tiles, (n_reads + top_reads - 1) / top_reads);
#define TOP(LTV)                                                                                        \
	do {                                                                                            \
		if (peak)                                                                               \
			cuda_emu::launch(g, dim3(kThreads), Large2TopSmem<LTV>::bytes, [&]() { large2_top_kernel<LTV, true>(p); }); \
		else                                                                                    \
			cuda_emu::launch(g, dim3(kThreads), Large2TopSmem<LTV>::bytes, [&]() { large2_top_kernel<LTV, false>(p); }); \
	} while (0)
	switch (lt) {
	case 1: TOP(1); break;
	case 2: TOP(2); break;
	case 3: TOP(3); break;
	case 4: TOP(4); break;
	default: TOP(5); break;
	}
#undef TOP
}

/* large path (bin_e 13..21) through the emulator: u8 reads [n_reads][2N] -> spectra.
 * in16 != 0: `reads` are decimated c16 images [n_reads][N] and `sums_in` their DC sums. */
void emu_large(int L, int peak, int in16, const uint8_t *reads, int n_reads, const int *hop_of,
	       const int *tw, const uint16_t *win, const long long *sums_in, long long *avg)
{
	const size_t N = (size_t)1 << L;
	if (L <= 17 && !g_emu_large3) {
		/* 5 CTAs share the (read, tile) items of the mid kernel in equal runs; 2 reads per top CTA */
		emu_large2(L, peak, in16, reads, n_reads, hop_of, (const int2 *)tw, win, sums_in, avg, 5, 2);
		return;
	}
	std::vector<long long> offs(n_reads);
	for (int i = 0; i < n_reads; i++)
		offs[i] = (long long)i * (long long)(in16 ? 4 * N : 2 * N);
	std::vector<c16> scratch((size_t)n_reads * N, 0xDEADBEEFu);
	std::vector<long long> sums((size_t)n_reads * 2, 0);
	LargeParams p;
	memset(&p, 0, sizeof(p));
	p.base = reads;
	p.read_off = offs.data();
	p.entry_base = 0;
	p.hop_of = hop_of;
	p.scratch = scratch.data();
	p.dc_sums = sums.data();
	std::vector<long long> smp(4096, 0);
	p.avg = avg;
	p.samples = smp.data();
	p.samples_per_read = 1;
	p.tw = (const int2 *)tw;
	std::vector<int2> twb((size_t)256 * 240, int2{ 0, 0 });
	{
		const int lbb = L - 8 < 8 ? L - 8 : 8;
		for (int se = 4; se < lbb; se++)
			for (int plow = 0; plow < 256; plow++)
				for (int ilow = 0; ilow < (1 << se); ilow++)
					twb[(size_t)256 * ((1 << se) - 16) + ((size_t)plow << se) + ilow] =
						p.tw[((size_t)((ilow << 8) | plow)) << (L - 9 - se)];
	}
	p.twb = twb.data();
	p.win = win;
	p.L = L;
	p.n_entries = n_reads;
	fill_tw0(p.tw0, p.tw, L);
	dim3 tiles((unsigned)(N / kWS), n_reads);
	if (!in16) {
		DcSumU8Params d;
		d.base = reads;
		d.read_off = offs.data();
		d.entry_base = 0;
		d.buf_len = (int)(2 * N);
		d.sums = sums.data();
		cuda_emu::launch(dim3(3, n_reads), dim3(256), 0, [&]() { dc_sums_u8_kernel(d); });
		cuda_emu::launch(tiles, dim3(kThreads), kLargeSmemA, [&]() { large_round_a_kernel<false>(p); });
	} else {
		memcpy(sums.data(), sums_in, sums.size() * 8);
		cuda_emu::launch(tiles, dim3(kThreads), kLargeSmemA, [&]() { large_round_a_kernel<true>(p); });
	}
	const int lb = L - 8 < 8 ? L - 8 : 8;
	const bool last = 8 + lb == L;
#define RB(LBV, LASTV)                                                                                  \
	do {                                                                                            \
		if (peak)                                                                               \
			cuda_emu::launch(tiles, dim3(kThreads), kLargeSmemB, [&]() { large_round_b_kernel<LBV, LASTV, true>(p); }); \
		else                                                                                    \
			cuda_emu::launch(tiles, dim3(kThreads), kLargeSmemB, [&]() { large_round_b_kernel<LBV, LASTV, false>(p); }); \
	} while (0)
	if (lb == 5) RB(5, true);
	else if (lb == 6) RB(6, true);
	else if (lb == 7) RB(7, true);
	else if (lb == 9) RB(9, true);
	else if (lb == 10) RB(10, true);
	else if (last) RB(8, true);
	else RB(8, false);
#undef RB
	if (8 + lb < L) {
		dim3 g(65536 / kThreads, (n_reads + kRoundCReads - 1) / kRoundCReads);
#define RC(LCV)                                                                                         \
	do {                                                                                            \
		if (peak)                                                                               \
			cuda_emu::launch(g, dim3(kThreads), 0, [&]() { large_round_c_kernel<LCV, true>(p); }); \
		else                                                                                    \
			cuda_emu::launch(g, dim3(kThreads), 0, [&]() { large_round_c_kernel<LCV, false>(p); }); \
	} while (0)
		switch (L - 16) {
		case 1: RC(1); break;
		case 2: RC(2); break;
		case 3: RC(3); break;
		case 4: RC(4); break;
		case 5: RC(5); break;
		}
#undef RC
	}
}

/* large path (bin_e 13..21) through the emulator: u8 reads [n_reads][2N] -> spectra.
 * in16 != 0: `reads` are decimated c16 images [n_reads][N] and `sums_in` their DC sums. */
static int g_large_pipe = 1;
void emu_set_large_pipe(int v) { g_large_pipe = v; }

void emu_large(int L, int peak, int in16, const uint8_t *reads, int n_reads, const int *hop_of,
	       const int *tw, const uint16_t *win, const long long *sums_in, long long *avg)
{
	const size_t N = (size_t)1 << L;
	std::vector<long long> offs(n_reads);
	for (int i = 0; i < n_reads; i++)
		offs[i] = (long long)i * (long long)(in16 ? 4 * N : 2 * N);
	std::vector<c16> scratch((size_t)n_reads * N, 0xDEADBEEFu);
	std::vector<long long> sums((size_t)n_reads * 2, 0);
	std::vector<unsigned> tickets(n_reads, 0);
	std::vector<int2> consts(n_reads, int2{ -1, -1 });
	std::vector<int2> twc_a(240);
	for (int st = 4; st < 8; st++)
		for (int m = 0; m < (1 << st); m++)
			twc_a[(size_t)(1 << st) - 16 + m] = ((const int2 *)tw)[(size_t)m << (L - 1 - st)];
	LargeParams p;
	memset(&p, 0, sizeof(p));
	p.base = reads;
	p.read_off = offs.data();
	p.entry_base = 0;
	p.hop_of = hop_of;
	p.scratch = scratch.data();
	p.dc_sums = sums.data();
	p.dc_consts = consts.data();
	p.twc_a = twc_a.data();
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
	p.tiles_log2 = L - 12;
	const unsigned pipe_grid = 5; /* runs that start and end inside a read */
	if (!in16) {
		DcSumU8Params d;
		d.base = reads;
		d.read_off = offs.data();
		d.entry_base = 0;
		d.buf_len = (int)(2 * N);
		d.sums = sums.data();
		d.tickets = tickets.data();
		d.consts = consts.data();
		cuda_emu::launch(dim3((L & 1) ? 1 : 3, n_reads), dim3(256), 0, [&]() { dc_sums_u8_kernel(d); }); /* both forms: one CTA per read / ticketed */
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
	else if (g_large_pipe)
		cuda_emu::launch(dim3(pipe_grid), dim3(kThreads), kLargeSmemBP, [&]() { large_round_b_pipe_kernel(p); });
	else RB(8, false);
#undef RB
	if (8 + lb < L) {
		const int lc = L - 16, cta_x = 65536 / (kThreads * round_c_vec(lc));
		p.c_reads = n_reads > 2 ? 2 : 1; /* more than one CTA row and a hop change inside a CTA's run */
		dim3 g(cta_x, (n_reads + p.c_reads - 1) / p.c_reads);
#define RC(LCV)                                                                                         \
	do {                                                                                            \
		if (peak)                                                                               \
			cuda_emu::launch(g, dim3(kThreads), 0, [&]() { large_round_c_kernel<LCV, true, round_c_vec(LCV)>(p); }); \
		else                                                                                    \
			cuda_emu::launch(g, dim3(kThreads), 0, [&]() { large_round_c_kernel<LCV, false, round_c_vec(LCV)>(p); }); \
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

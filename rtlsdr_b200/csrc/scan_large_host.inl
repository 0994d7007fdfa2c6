/* Host side of the N >= 8192 path (see scan_large.cuh).  Included by scan_abi.cu
 * after the handle definition. */
namespace {

/* launch programmatically dependent on the previous kernel in the stream (the kernel has a pdl_wait()) */
template <class K, class P>
int launch_pdl(rtlsdr_gpu_scan *h, K kern, dim3 grid, int smem, const P &prm)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = dim3(kThreads);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = h->stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	CU(cudaLaunchKernelEx(&cfg, kern, prm));
	return 0;
}

template <int LB, bool LAST>
int launch_round_b_t(rtlsdr_gpu_scan *h, const LargeParams &p, dim3 grid)
{
	const int smem = kLargeSmemB;
	if (h->cfg.peak_hold) {
		auto k = large_round_b_kernel<LB, LAST, true>;
		CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		k<<<grid, kThreads, smem, h->stream>>>(p);
	} else {
		auto k = large_round_b_kernel<LB, LAST, false>;
		CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		k<<<grid, kThreads, smem, h->stream>>>(p);
	}
	return check_launch(h, "large_round_b_kernel");
}

int launch_round_b(rtlsdr_gpu_scan *h, const LargeParams &p, dim3 grid, int lb, bool last)
{
	switch (lb) {
	case 5: return last ? launch_round_b_t<5, true>(h, p, grid) : RTLSDR_GPU_ERR_CONFIG;
	case 6: return last ? launch_round_b_t<6, true>(h, p, grid) : RTLSDR_GPU_ERR_CONFIG;
	case 7: return last ? launch_round_b_t<7, true>(h, p, grid) : RTLSDR_GPU_ERR_CONFIG;
	case 8: return last ? launch_round_b_t<8, true>(h, p, grid) : launch_round_b_t<8, false>(h, p, grid);
	case 9: return last ? launch_round_b_t<9, true>(h, p, grid) : RTLSDR_GPU_ERR_CONFIG;
	case 10: return last ? launch_round_b_t<10, true>(h, p, grid) : RTLSDR_GPU_ERR_CONFIG;
	default: return RTLSDR_GPU_ERR_CONFIG;
	}
}

template <int LC>
int launch_round_c_t(rtlsdr_gpu_scan *h, const LargeParams &p, dim3 grid)
{
	int rc;
	if (h->cfg.peak_hold)
		rc = launch_pdl(h, large_round_c_kernel<LC, true, round_c_vec(LC)>, grid, 0, p);
	else
		rc = launch_pdl(h, large_round_c_kernel<LC, false, round_c_vec(LC)>, grid, 0, p);
	if (rc)
		return rc;
	return check_launch(h, "large_round_c_kernel");
}

int launch_round_c(rtlsdr_gpu_scan *h, const LargeParams &p, dim3 grid, int lc)
{
	switch (lc) {
	case 1: return launch_round_c_t<1>(h, p, grid);
	case 2: return launch_round_c_t<2>(h, p, grid);
	case 3: return launch_round_c_t<3>(h, p, grid);
	case 4: return launch_round_c_t<4>(h, p, grid);
	case 5: return launch_round_c_t<5>(h, p, grid);
	default: return RTLSDR_GPU_ERR_CONFIG;
	}
}

/*
 * Entries [0, n_reads) of the (hop-sorted) batch, `chunk` at a time.
 * Scratch per chunk: [data chunk x N c16][sums chunk x 2 int64][decimation scratch].
 */
int large_process(rtlsdr_gpu_scan *h, const uint8_t *base, const long long *d_offs, const int *d_hops, int n_reads)
{
	const int L = h->cfg.bin_e;
	const size_t N = (size_t)1 << L;
	const bool decim = (h->cfg.boxcar && h->cfg.downsample > 1) || h->cfg.downsample_passes > 0;
	const size_t per = N * 4 + 32 + (decim ? DecimScratch::per_entry(h) : 0); /* data, sums, tickets + constants */
	const size_t extra = decim ? DecimScratch::slack() + 512 : 512;
	const int chunk_max = (int)std::max<size_t>(1, kScratchBudget / per);
	int rc;
	for (int e0 = 0; e0 < n_reads; e0 += chunk_max) {
		const int cnt = std::min(chunk_max, n_reads - e0);
		if ((rc = ensure_scratch(h, per * (size_t)cnt + extra)))
			return rc;
		uint8_t *sp = h->d_scratch;
		c16 *data = (c16 *)sp;
		sp += (size_t)cnt * N * 4;
		long long *sums = (long long *)sp;
		sp += (size_t)cnt * 16;
		unsigned *tickets = (unsigned *)sp;   /* directly behind the sums: one memset clears both */
		sp += (size_t)cnt * 8;
		int2 *consts = (int2 *)sp;
		sp += (size_t)cnt * 8;

		LargeParams p;
		memset(&p, 0, sizeof(p));
		p.entry_base = e0;
		p.hop_of = d_hops;
		p.scratch = data;
		p.dc_sums = sums;
		p.dc_consts = consts;
		p.twc_a = h->d_twc;
		p.avg = h->d_avg;
		p.samples = h->d_smp64;
		p.samples_per_read = h->samples_per_read;
		p.tw = h->d_tw;
		p.twb = h->d_twb;
		p.win = h->d_win;
		p.L = L;
		p.n_entries = cnt;
		p.tw0 = h->tw0;

		const int smem_a = kLargeSmemA;
		dim3 grid_tiles((unsigned)(N / kWS), (unsigned)cnt);
		/* pipelined round B: one resident wave of CTAs, equal runs of the (tile, read) items */
		p.tiles_log2 = L - 12;
		const unsigned grid_pipe = (unsigned)std::min<long long>((long long)cnt << (L - 12), 2ll * h->num_sms);
		if (!decim) {
			/* 32 KiB per CTA, up to 32 CTAs per read (one CTA per read took 30 us instead of 16-22 for 256 reads of
			 * 256 KiB: too few bytes in flight) */
			const unsigned per_read =
				(unsigned)std::max<size_t>(1, std::min<size_t>(32, (size_t)h->cfg.buf_len / (256 * 16 * 8)));
			if (per_read > 1)
				CU(cudaMemsetAsync(sums, 0, (size_t)cnt * 24, h->stream));
			DcSumU8Params d;
			d.base = base;
			d.read_off = d_offs;
			d.entry_base = e0;
			d.buf_len = h->cfg.buf_len;
			d.sums = sums;
			d.tickets = tickets;
			d.consts = consts;
			dim3 g(per_read, (unsigned)cnt);
			dc_sums_u8_kernel<<<g, 256, 0, h->stream>>>(d);
			if ((rc = check_launch(h, "dc_sums_u8_kernel")))
				return rc;
			p.base = base;
			p.read_off = d_offs;
			auto k = large_round_a_kernel<false>;
			CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
			if ((rc = launch_pdl(h, k, grid_tiles, smem_a, p)))
				return rc;
			if ((rc = check_launch(h, "large_round_a_kernel")))
				return rc;
		} else {
			/* decimate into c16 images placed after the sums, then transform from them */
			DecimScratch sc(h, cnt, (uint8_t *)(((uintptr_t)sp + 255) & ~(uintptr_t)255));
			if ((rc = run_decimators(h, base, d_offs + e0, cnt, sc)))
				return rc;
			CU(cudaMemcpyAsync(sums, sc.sums, (size_t)cnt * 16, cudaMemcpyDeviceToDevice, h->stream));
			p.base = (const uint8_t *)sc.img;
			p.read_off = nullptr;
			p.regular_stride = h->image_stride * 4;
			auto k = large_round_a_kernel<true>;
			CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
			k<<<grid_tiles, kThreads, smem_a, h->stream>>>(p);
			if ((rc = check_launch(h, "large_round_a_kernel")))
				return rc;
		}
		/* (letting round B take 9-10 stages to save round C was measured slower: 32-byte runs) */
		const int lb = std::min(8, L - 8);
		if (8 + lb < L && h->dbg_large_pipe) {
			auto k = large_round_b_pipe_kernel;
			CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kLargeSmemBP));
			if ((rc = launch_pdl(h, k, dim3(grid_pipe), kLargeSmemBP, p)))
				return rc;
			if ((rc = check_launch(h, "large_round_b_pipe_kernel")))
				return rc;
		} else if ((rc = launch_round_b(h, p, grid_tiles, lb, 8 + lb == L)))
			return rc;
		if (8 + lb < L) {
			const int cta_x = 65536 / (kThreads * round_c_vec(L - 16));
			p.c_reads = round_c_reads(cnt, cta_x);
			dim3 g((unsigned)cta_x, (unsigned)((cnt + p.c_reads - 1) / p.c_reads));
			if ((rc = launch_round_c(h, p, g, L - 16)))
				return rc;
		}
	}
	return 0;
}

} // namespace

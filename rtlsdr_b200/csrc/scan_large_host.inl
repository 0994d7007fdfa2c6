namespace {
int large_process(rtlsdr_gpu_scan *h, const uint8_t *, const long long *, const int *, int)
{
	h->last_error = "large FFT path not built";
	return RTLSDR_GPU_ERR_CONFIG;
}
} // namespace

/*
 * TEST INFRASTRUCTURE -- not part of the shipped GPU path.
 *
 * Housekeeping half of the fake librtlsdr the unmodified reference objects
 * (src/rtl_power.c, src/convenience/convenience.c) link against.  The
 * streaming half (read_sync / read_async / centre frequency) lives in
 * host/synth_source.c.  These calls program dongle hardware in the reference
 * (src/librtlsdr.c) and have no effect on the scan arithmetic, so they succeed
 * and do nothing.
 */
#include <stdint.h>

typedef struct rtlsdr_dev rtlsdr_dev_t;

int rtlsdr_get_tuner_gains(rtlsdr_dev_t *dev, int *gains)
{
	static const int table[] = { 0, 9, 14, 27, 37, 77, 87, 125, 144, 157, 166, 197 };
	int i, n = (int)(sizeof(table) / sizeof(table[0]));
	(void)dev;
	if (gains)
		for (i = 0; i < n; i++)
			gains[i] = table[i];
	return n;
}
int rtlsdr_set_tuner_gain(rtlsdr_dev_t *dev, int gain) { (void)dev; (void)gain; return 0; }
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *dev, int manual) { (void)dev; (void)manual; return 0; }
int rtlsdr_set_and_get_tuner_bandwidth(rtlsdr_dev_t *dev, uint32_t bw, uint32_t *applied_bw, int apply_bw)
{
	(void)dev; (void)apply_bw;
	if (applied_bw)
		*applied_bw = bw;
	return 0;
}
int rtlsdr_set_direct_sampling(rtlsdr_dev_t *dev, int on) { (void)dev; (void)on; return 0; }
int rtlsdr_set_ds_mode(rtlsdr_dev_t *dev, int mode, uint32_t freq_threshold)
{
	(void)dev; (void)mode; (void)freq_threshold;
	return 0;
}
int rtlsdr_set_offset_tuning(rtlsdr_dev_t *dev, int on) { (void)dev; (void)on; return 0; }
int rtlsdr_set_freq_correction_ppb(rtlsdr_dev_t *dev, int ppb) { (void)dev; (void)ppb; return 0; }
int rtlsdr_set_bias_tee(rtlsdr_dev_t *dev, int on) { (void)dev; (void)on; return 0; }
const char *rtlsdr_get_ver_id(void) { return "synthetic"; }
uint32_t rtlsdr_get_version(void) { return 0; }

/*
 * TEST INFRASTRUCTURE -- not part of the shipped GPU path.
 *
 * scan_oracle: plain-C restatement of rtl_power's per-hop scan arithmetic
 * (reference /root/reference/src/rtl_power.c:247-327, 329-436, 554-765).
 * It is the checker the CUDA path is compared against where the compiled
 * reference (oracle/_ref) cannot travel or where a per-read entry point is
 * needed.  Pinned: tests/test_oracle_vs_ref.py compares every function here
 * against the reference object itself and against tests/golden/ (generated
 * from the reference by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use this file; the product library never links or calls it.
 */
#ifndef SCAN_ORACLE_H
#define SCAN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int bin_e;              /* log2 of the FFT length (0 = rms path) */
	int buf_len;            /* bytes per read */
	int downsample;         /* ds */
	int downsample_passes;  /* ds_p */
	int boxcar;             /* 1 unless -F given */
	int comp_fir_size;      /* -F argument */
	int peak_hold;          /* -P */
	const int32_t *window;  /* [1 << bin_e] host-built coefficients */
	const int16_t *sine;    /* [3N/4] host-built sine table, N_WAVE = 1 << bin_e */
} oracle_cfg_t;

/* rtl_power.c:247-261 */
void oracle_sine_table(int m, int16_t *out);
/* rtl_power.c:329-408 + :985-988; returns 0, or -1 for an unknown name (table then = rectangle) */
int oracle_window_coefs(const char *name, int n, int32_t *out);
/* rtl_power.c:263-269 */
int16_t oracle_fix_mpy(int16_t a, int16_t b);
/* rtl_power.c:271-327; sine table must have been built for log2_nwave >= m */
int oracle_fix_fft(int16_t *iq, int m, const int16_t *sine, int log2_nwave);
/* rtl_power.c:554-579 */
void oracle_fifth_order(int16_t *data, int length);
/* rtl_power.c:598-626; fir = {9, c1..c9} row of cic_9_tables */
void oracle_generic_fir(int16_t *data, int length, const int *fir);
/* rtl_power.c:219-232 */
const int *oracle_cic9(int passes);
/* rtl_power.c:581-596 */
void oracle_remove_dc(int16_t *data, int length);
/* rtl_power.c:410-436 */
int64_t oracle_rms_power(const uint8_t *buf, int buf_len, int64_t avg0, int peak_hold);
/*
 * One hop visit = the body of scanner()'s loop after the read
 * (rtl_power.c:660-718).  `work` is scratch of buf_len int16.
 * avg[1 << bin_e] and *samples are updated in place.
 */
void oracle_scan_read(const oracle_cfg_t *cfg, const uint8_t *buf8, int16_t *work,
		      int64_t *avg, int *samples);
/*
 * Numeric half of csv_dbm() (rtl_power.c:722-765): DC nuke + half swap in
 * place on avg[], then dB for bins i1..i2 plus the re-associated duplicate
 * of the last bin.  Returns the number of doubles written to db (i2-i1+2).
 * Does not zero avg (the caller decides), unlike the reference.
 */
int oracle_epilogue(int64_t *avg, int bin_e, double crop, int rate, int samples, double *db);

#ifdef __cplusplus
}
#endif
#endif

/*
 * TEST INFRASTRUCTURE -- not part of the shipped GPU path.
 *
 * ref_harness: drives the UNMODIFIED reference object
 * (/root/reference/src/rtl_power.c compiled with -Dmain=rtl_power_main
 * -Dusleep=ref_usleep_noop, see oracle/Makefile) the way its own main() does
 * (rtl_power.c:894 frequency_range, :978-988 sample rate / sine_table /
 * fft_buf / window_coefs, :989-990 scanner loop, :995-1000 csv_dbm) and
 * exposes the results through a flat C interface for ctypes.
 *
 * Nothing in here re-implements DSP: every number comes out of the reference's
 * own scanner()/fix_fft()/csv_dbm() code.  The sample source is
 * host/synth_source.c (fake librtlsdr).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../host/synth_source.h"

/* ---- the reference's externally visible state (rtl_power.c:85-120) ---- */
struct tuning_state {
	int freq;
	int rate;
	int bin_e;
	long *avg;
	int samples;
	int downsample;
	int downsample_passes;
	double crop;
	uint8_t *buf8;
	int buf_len;
};
extern struct tuning_state tunes[];
extern int tune_count;
extern int boxcar, comp_fir_size, peak_hold;
extern int16_t *Sinewave;
extern double *power_table;
extern int N_WAVE, LOG2_N_WAVE;
extern int16_t *fft_buf;
extern int *window_coefs;
extern FILE *file;
extern int cic_9_tables[][10];

void frequency_range(char *arg, double crop);
void sine_table(int size);
int fix_fft(int16_t iq[], int m);
void scanner(void);
void csv_dbm(struct tuning_state *ts);
void fifth_order(int16_t *data, int length);
void generic_fir(int16_t *data, int length, int *fir);
void remove_dc(int16_t *data, int length);
void downsample_iq(int16_t *data, int length);
void rms_power(struct tuning_state *ts);
long real_conj(int16_t real, int16_t imag);
double rectangle(int, int);
double hamming(int, int);
double blackman(int, int);
double blackman_harris(int, int);
double hann_poisson(int, int);
double youssef(int, int);
double kaiser(int, int);
double bartlett(int, int);

/* retune() sleeps 5 ms per hop change (rtl_power.c:548); compiled away */
int ref_usleep_noop(unsigned int usec)
{
	(void)usec;
	return 0;
}

typedef double (*window_fn_t)(int, int);

static window_fn_t window_by_name(const char *name)
{
	/* same table as the -w option parser, rtl_power.c:826-843; unknown
	 * names silently stay rectangle there, so they do here */
	static const struct { const char *n; window_fn_t f; } tab[] = {
		{ "rectangle", rectangle }, { "hamming", hamming }, { "blackman", blackman },
		{ "blackman-harris", blackman_harris }, { "hann-poisson", hann_poisson },
		{ "youssef", youssef }, { "kaiser", kaiser }, { "bartlett", bartlett },
	};
	size_t i;
	for (i = 0; name && i < sizeof(tab) / sizeof(tab[0]); i++)
		if (strcmp(tab[i].n, name) == 0)
			return tab[i].f;
	return rectangle;
}

static void drop_state(void)
{
	int i;
	for (i = 0; i < tune_count; i++) {
		free(tunes[i].avg);
		free(tunes[i].buf8);
		tunes[i].avg = NULL;
		tunes[i].buf8 = NULL;
	}
	tune_count = 0;
	free(Sinewave);     Sinewave = NULL;
	free(power_table);  power_table = NULL;
	free(fft_buf);      fft_buf = NULL;
	free(window_coefs); window_coefs = NULL;
}

/*
 * Same sequence as main(): option globals, frequency_range(), sine_table(),
 * fft_buf, window_coefs.  fir_arg < 0 means "-F not given"; otherwise -F <fir_arg>.
 * Returns tune_count.
 */
int ref_configure(const char *freq_arg, double crop, const char *window, int fir_arg, int peak)
{
	char *arg;
	int i, length, freqs[3000];
	window_fn_t fn = window_by_name(window);
	FILE *saved;

	drop_state();
	boxcar = 1;
	comp_fir_size = 0;
	if (fir_arg >= 0) {
		boxcar = 0;
		comp_fir_size = fir_arg;
	}
	peak_hold = peak ? 1 : 0;

	arg = strdup(freq_arg);
	/* the planner's report goes to stderr; keep test logs quiet */
	saved = stderr;
	stderr = fopen("/dev/null", "w");
	frequency_range(arg, crop);
	if (stderr)
		fclose(stderr);
	stderr = saved;
	free(arg);
	if (tune_count <= 0)
		return tune_count;

	rtlsdr_set_sample_rate(NULL, (uint32_t)tunes[0].rate);
	sine_table(tunes[0].bin_e);
	fft_buf = malloc((size_t)tunes[0].buf_len * sizeof(int16_t));
	length = 1 << tunes[0].bin_e;
	window_coefs = malloc((size_t)length * sizeof(int));
	for (i = 0; i < length; i++)
		window_coefs[i] = (int)(256 * fn(i, length));

	for (i = 0; i < tune_count; i++)
		freqs[i] = tunes[i].freq;
	synth_set_hops(NULL, freqs, tune_count);
	synth_set_block_len(NULL, (size_t)tunes[0].buf_len);
	return tune_count;
}

/* out[0..7] = tune_count, bin_e, buf_len, downsample, downsample_passes, rate, boxcar, comp_fir_size */
void ref_plan(int *out)
{
	out[0] = tune_count;
	out[1] = tunes[0].bin_e;
	out[2] = tunes[0].buf_len;
	out[3] = tunes[0].downsample;
	out[4] = tunes[0].downsample_passes;
	out[5] = tunes[0].rate;
	out[6] = boxcar;
	out[7] = comp_fir_size;
}

double ref_crop(void) { return tunes[0].crop; }
int ref_hop_freq(int hop) { return tunes[hop].freq; }

void ref_source(int mode, uint64_t seed, int param)
{
	synth_configure(NULL, mode, seed, param);
}

void ref_source_replay(const uint8_t *pool, size_t read_len, size_t n_reads)
{
	synth_set_replay(NULL, pool, read_len, n_reads);
}

/* run `passes` full sweeps exactly like the main loop body (rtl_power.c:989-990) */
void ref_scan(int passes)
{
	int i;
	for (i = 0; i < passes; i++)
		scanner();
}

/* wall-clock seconds for `passes` sweeps; used by bench.py's CPU baseline legs */
double ref_scan_timed(int passes)
{
	struct timespec a, b;
	clock_gettime(CLOCK_MONOTONIC, &a);
	ref_scan(passes);
	clock_gettime(CLOCK_MONOTONIC, &b);
	return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

int ref_samples(int hop) { return tunes[hop].samples; }

void ref_avg(int hop, int64_t *out)
{
	int i, n = 1 << tunes[hop].bin_e;
	for (i = 0; i < n; i++)
		out[i] = (int64_t)tunes[hop].avg[i];
}

void ref_window_coefs(int32_t *out)
{
	int i, n = 1 << tunes[0].bin_e;
	for (i = 0; i < n; i++)
		out[i] = window_coefs[i];
}

void ref_sinewave(int16_t *out)
{
	int i;
	for (i = 0; i < N_WAVE * 3 / 4; i++)
		out[i] = Sinewave[i];
}

/* FNV-1a over every int64 bin of every hop, natural bin order */
uint64_t ref_fnv(void)
{
	uint64_t h = 14695981039346656037ULL;
	int t, i;
	for (t = 0; t < tune_count; t++) {
		int n = 1 << tunes[t].bin_e;
		for (i = 0; i < n; i++) {
			h ^= (uint64_t)tunes[t].avg[i];
			h *= 1099511628211ULL;
		}
	}
	return h;
}

/* csv_dbm() for one hop into a caller buffer; like the reference this also
 * rewrites avg[] (DC nuke + half swap) and then zeroes avg[]/samples.
 * Returns the number of characters (excluding NUL) or -1. */
int ref_csv(int hop, char *out, int cap)
{
	char *mem = NULL;
	size_t sz = 0;
	int n;
	file = open_memstream(&mem, &sz);
	if (!file)
		return -1;
	csv_dbm(&tunes[hop]);
	fclose(file);
	file = NULL;
	n = (int)sz;
	if (n >= cap)
		n = cap - 1;
	memcpy(out, mem, (size_t)n);
	out[n] = '\0';
	free(mem);
	return (int)sz;
}

/* ---- unit-level entry points into the reference object ----------------- */

void ref_sine_table(int m)
{
	free(Sinewave);
	free(power_table);
	sine_table(m);
}
int ref_fix_fft(int16_t *iq, int m) { return fix_fft(iq, m); }
void ref_fifth_order(int16_t *data, int length) { fifth_order(data, length); }
void ref_downsample_iq(int16_t *data, int length) { downsample_iq(data, length); }
void ref_generic_fir(int16_t *data, int length, int table) { generic_fir(data, length, cic_9_tables[table]); }
void ref_cic9(int table, int *out) { memcpy(out, cic_9_tables[table], 10 * sizeof(int)); }
void ref_remove_dc(int16_t *data, int length) { remove_dc(data, length); }
long ref_real_conj(int16_t re, int16_t im) { return real_conj(re, im); }
double ref_window(const char *name, int i, int length) { return window_by_name(name)(i, length); }

/* rms_power() on caller bytes with a private tuning_state (rtl_power.c:410-436) */
long ref_rms_power(const uint8_t *buf, int buf_len, long avg0, int peak)
{
	struct tuning_state ts;
	long avg = avg0;
	int saved = peak_hold;
	memset(&ts, 0, sizeof(ts));
	ts.buf8 = (uint8_t *)buf;
	ts.buf_len = buf_len;
	ts.avg = &avg;
	peak_hold = peak;
	rms_power(&ts);
	peak_hold = saved;
	return avg;
}

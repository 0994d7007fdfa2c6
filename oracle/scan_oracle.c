/*
 * TEST INFRASTRUCTURE -- not part of the shipped GPU path.  See scan_oracle.h.
 *
 * Each function restates one piece of /root/reference/src/rtl_power.c in the
 * closed forms of SURVEY.md section 8(a); the reference lines are cited per
 * function.  All integer arithmetic wraps exactly like the reference built
 * with gcc (int16 stores are modulo 2^16, int sums modulo 2^32).
 */
#include "scan_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int16_t wrap16(int32_t v)
{
	return (int16_t)(uint16_t)((uint32_t)v & 0xFFFFu);
}

/* ---- tables ----------------------------------------------------------- */

/* rtl_power.c:247-261: Sinewave[i] = (int)round(32767*sin(2*pi*i/N)), i < 3N/4 */
void oracle_sine_table(int m, int16_t *out)
{
	int n = 1 << m, count = n * 3 / 4, i;
	for (i = 0; i < count; i++) {
		double d = (double)i * 2.0 * M_PI / n;
		out[i] = (int16_t)(int)round(32767 * sin(d));
	}
}

/* window shapes, rtl_power.c:329-408 (argument order and operation order kept,
 * since the truncation to int below is sensitive to the last ulp) */
static double w_rectangle(int i, int n) { (void)i; (void)n; return 1.0; }

static double w_hamming(int i, int n)
{
	double a = 25.0 / 46.0, b = 21.0 / 46.0, n1 = (double)(n - 1);
	return a - b * cos(2 * i * M_PI / n1);
}

static double w_blackman(int i, int n)
{
	double a0 = 7938.0 / 18608.0, a1 = 9240.0 / 18608.0, a2 = 1430.0 / 18608.0;
	double n1 = (double)(n - 1);
	return a0 - a1 * cos(2 * i * M_PI / n1) + a2 * cos(4 * i * M_PI / n1);
}

static double w_blackman_harris(int i, int n)
{
	double a0 = 0.35875, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
	double n1 = (double)(n - 1);
	return a0 - a1 * cos(2 * i * M_PI / n1) + a2 * cos(4 * i * M_PI / n1)
		  - a3 * cos(6 * i * M_PI / n1);
}

static double w_hann_poisson(int i, int n)
{
	double a = 2.0, n1 = (double)(n - 1);
	return 0.5 * (1 - cos(2 * M_PI * i / n1)) *
	       pow(M_E, (-a * (double)abs((int)(n1 - 1 - 2 * i))) / n1);
}

static double w_youssef(int i, int n)
{
	double a = 0.0025, n1 = (double)(n - 1);
	double w = w_blackman_harris(i, n);
	w *= pow(M_E, (-a * (double)abs((int)(n1 - 1 - 2 * i))) / n1);
	return w;
}

static double w_bartlett(int i, int n)
{
	double l = (double)n, n1 = l - 1;
	double w = (i - n1 / 2) / (l / 2);
	if (w < 0)
		w = -w;
	return 1 - w;
}

/* rtl_power.c:826-843 name table (kaiser is rectangle, :392-396) + :985-988 */
int oracle_window_coefs(const char *name, int n, int32_t *out)
{
	static const struct { const char *name; double (*fn)(int, int); } tab[] = {
		{ "rectangle", w_rectangle }, { "hamming", w_hamming },
		{ "blackman", w_blackman }, { "blackman-harris", w_blackman_harris },
		{ "hann-poisson", w_hann_poisson }, { "youssef", w_youssef },
		{ "kaiser", w_rectangle }, { "bartlett", w_bartlett },
	};
	double (*fn)(int, int) = w_rectangle;
	int i, rc = -1;
	for (i = 0; name && i < (int)(sizeof(tab) / sizeof(tab[0])); i++) {
		if (strcmp(name, tab[i].name) == 0) {
			fn = tab[i].fn;
			rc = 0;
		}
	}
	for (i = 0; i < n; i++)
		out[i] = (int32_t)(256 * fn(i, n));
	return rc;
}

/* rtl_power.c:219-232 (values are data, scaled 2^15; row = downsample_passes) */
static const int cic9_rows[11][10] = {
	{ 0 },
	{ 9, -156, -97, 2798, -15489, 61019, -15489, 2798, -97, -156 },
	{ 9, -128, -568, 5593, -24125, 74126, -24125, 5593, -568, -128 },
	{ 9, -129, -639, 6187, -26281, 77511, -26281, 6187, -639, -129 },
	{ 9, -122, -612, 6082, -26353, 77818, -26353, 6082, -612, -122 },
	{ 9, -120, -602, 6015, -26269, 77757, -26269, 6015, -602, -120 },
	{ 9, -120, -582, 5951, -26128, 77542, -26128, 5951, -582, -120 },
	{ 9, -119, -580, 5931, -26094, 77505, -26094, 5931, -580, -119 },
	{ 9, -119, -578, 5921, -26077, 77484, -26077, 5921, -578, -119 },
	{ 9, -119, -577, 5917, -26067, 77473, -26067, 5917, -577, -119 },
	{ 9, -199, -362, 5303, -25505, 77489, -25505, 5303, -362, -199 },
};

const int *oracle_cic9(int passes)
{
	if (passes < 0 || passes > 10)
		return NULL;
	return cic9_rows[passes];
}

/* ---- fixed-point FFT -------------------------------------------------- */

/* rtl_power.c:263-269: ((a*b)>>14, then (c>>1)+(c&1)) == (a*b + 2^14) >> 15 */
int16_t oracle_fix_mpy(int16_t a, int16_t b)
{
	int32_t p = (int32_t)a * (int32_t)b;
	return wrap16((p + 16384) >> 15);
}

static unsigned bit_reverse(unsigned v, int bits)
{
	unsigned r = 0;
	int i;
	for (i = 0; i < bits; i++) {
		r = (r << 1) | (v & 1u);
		v >>= 1;
	}
	return r;
}

/* rtl_power.c:271-327: radix-2 DIT, halving on every stage, halved twiddles */
int oracle_fix_fft(int16_t *iq, int m, const int16_t *sine, int log2_nwave)
{
	int n = 1 << m, nwave = 1 << log2_nwave;
	int s, g, i;
	if (n > nwave)
		return -1;
	/* :282-297 bit-reversal permutation, each pair swapped once */
	for (i = 1; i < n; i++) {
		int r = (int)bit_reverse((unsigned)i, m);
		if (r > i) {
			int16_t tr = iq[2 * i], ti = iq[2 * i + 1];
			iq[2 * i] = iq[2 * r];
			iq[2 * i + 1] = iq[2 * r + 1];
			iq[2 * r] = tr;
			iq[2 * r + 1] = ti;
		}
	}
	/* :298-324 */
	for (s = 0; s < m; s++) {
		int half = 1 << s, span = half << 1;
		int k = log2_nwave - 1 - s;
		for (g = 0; g < half; g++) {
			int j = g << k;
			int16_t wr = (int16_t)(sine[j + nwave / 4] >> 1);
			int16_t wi = (int16_t)(wrap16(-(int32_t)sine[j]) >> 1);
			for (i = g; i < n; i += span) {
				int p = i + half;
				int16_t br = iq[2 * p], bi = iq[2 * p + 1];
				int16_t tr = wrap16((int32_t)oracle_fix_mpy(wr, br) - oracle_fix_mpy(wi, bi));
				int16_t ti = wrap16((int32_t)oracle_fix_mpy(wr, bi) + oracle_fix_mpy(wi, br));
				int16_t qr = (int16_t)(iq[2 * i] >> 1);
				int16_t qi = (int16_t)(iq[2 * i + 1] >> 1);
				iq[2 * p] = wrap16((int32_t)qr - tr);
				iq[2 * p + 1] = wrap16((int32_t)qi - ti);
				iq[2 * i] = wrap16((int32_t)qr + tr);
				iq[2 * i + 1] = wrap16((int32_t)qi + ti);
			}
		}
	}
	return 0;
}

/* ---- decimators ------------------------------------------------------- */

/*
 * rtl_power.c:554-579, one half (every 2nd int16) of an interleaved buffer.
 * Output n lands on int16 index 2n; with s[n] = data[2n] of the INPUT:
 *   n=0: ((s0+s1)*10 + (s2+s3)*5 + s3 + s5) >> 4
 *   n=1: ((s1+s2)*10 + (s0+s3)*5 + s4 + s5) >> 4
 *   n=2: (s0 + (s1+s4)*5 + (s2+s3)*10 + s5) >> 4
 *   n=3: (s2 + (s3+s5)*5 + (s4+s5)*10 + s6) >> 4        (s5 twice, :571-576)
 *   n=4: (s4 + (s5+s7)*5 + (s5+s6)*10 + s8) >> 4
 *   n>=5: (s[2n-5] + (s[2n-4]+s[2n-1])*5 + (s[2n-3]+s[2n-2])*10 + s[2n]) >> 4
 * for every n with 4n < length.  Reads always see input values because the
 * write index trails the read index, so a snapshot of the input is exact.
 */
void oracle_fifth_order(int16_t *data, int length)
{
	int count = (length + 1) / 2; /* samples of this half inside `length` */
	int32_t *s = (int32_t *)malloc((size_t)(count > 6 ? count : 6) * sizeof(int32_t));
	int n;
	for (n = 0; n < count; n++)
		s[n] = data[2 * n];
	data[0] = wrap16(((s[0] + s[1]) * 10 + (s[2] + s[3]) * 5 + s[3] + s[5]) >> 4);
	data[2] = wrap16(((s[1] + s[2]) * 10 + (s[0] + s[3]) * 5 + s[4] + s[5]) >> 4);
	data[4] = wrap16((s[0] + (s[1] + s[4]) * 5 + (s[2] + s[3]) * 10 + s[5]) >> 4);
	for (n = 3; 4 * n < length; n++) {
		int32_t a, b, c, d, e, f;
		if (n == 3) {
			a = s[2]; b = s[3]; c = s[4]; d = s[5]; e = s[5]; f = s[6];
		} else if (n == 4) {
			a = s[4]; b = s[5]; c = s[5]; d = s[6]; e = s[7]; f = s[8];
		} else {
			a = s[2 * n - 5]; b = s[2 * n - 4]; c = s[2 * n - 3];
			d = s[2 * n - 2]; e = s[2 * n - 1]; f = s[2 * n];
		}
		data[2 * n] = wrap16((a + (b + e) * 5 + (c + d) * 10 + f) >> 4);
	}
	free(s);
}

/*
 * rtl_power.c:598-626: samples 0..8 pass through; for k >= 9
 *   out[k] = ((h0+h8)*f1 + (h1+h7)*f2 + (h2+h6)*f3 + (h3+h5)*f4 + h4*f5) >> 15
 * with h[t] = in[k-9+t], the nine INPUT samples before k.  The int sum wraps.
 */
void oracle_generic_fir(int16_t *data, int length, const int *fir)
{
	int count = (length + 1) / 2;
	int32_t *s = (int32_t *)malloc((size_t)(count > 9 ? count : 9) * sizeof(int32_t));
	int k;
	for (k = 0; k < count; k++)
		s[k] = data[2 * k];
	for (k = 9; 2 * k < length; k++) {
		const int32_t *h = s + (k - 9);
		uint32_t sum = 0;
		sum += (uint32_t)(h[0] + h[8]) * (uint32_t)fir[1];
		sum += (uint32_t)(h[1] + h[7]) * (uint32_t)fir[2];
		sum += (uint32_t)(h[2] + h[6]) * (uint32_t)fir[3];
		sum += (uint32_t)(h[3] + h[5]) * (uint32_t)fir[4];
		sum += (uint32_t)h[4] * (uint32_t)fir[5];
		data[2 * k] = wrap16((int32_t)sum >> 15);
	}
	free(s);
}

/* rtl_power.c:581-596: mean over every 2nd element, divided by the
 * INTERLEAVED length, C division (toward zero), int16; early out on 0 */
void oracle_remove_dc(int16_t *data, int length)
{
	int64_t sum = 0;
	int16_t ave;
	int i;
	for (i = 0; i < length; i += 2)
		sum += data[i];
	ave = wrap16((int32_t)(sum / (int64_t)length));
	if (ave == 0)
		return;
	for (i = 0; i < length; i += 2)
		data[i] = wrap16((int32_t)data[i] - ave);
}

/* rtl_power.c:410-436 */
int64_t oracle_rms_power(const uint8_t *buf, int buf_len, int64_t avg0, int peak_hold)
{
	int64_t p = 0, t = 0;
	double dc, err;
	int i;
	for (i = 0; i < buf_len; i++) {
		int s = (int)buf[i] - 127;
		t += s;
		p += (int64_t)(s * s);
	}
	dc = (double)t / (double)buf_len;
	err = (double)(t * 2) * dc - dc * dc * (double)buf_len;
	p -= (int64_t)round(err);
	if (!peak_hold)
		return avg0 + p;
	return avg0 > p ? avg0 : p;
}

/* ---- one hop visit ---------------------------------------------------- */

/* rtl_power.c:671-681 in closed form: D[k] = wrap16(sum of the ds inputs
 * k*ds .. k*ds+ds-1 that exist), every other slot of the buffer becomes 0 */
static void boxcar_decimate(int16_t *work, int buf_len, int ds)
{
	int pairs = buf_len / 2, outs = (pairs + ds - 1) / ds, k, i;
	for (k = 0; k < outs; k++) {
		int32_t si = 0, sq = 0;
		for (i = 0; i < ds && k * ds + i < pairs; i++) {
			si += work[2 * (k * ds + i)];
			sq += work[2 * (k * ds + i) + 1];
		}
		/* slot k is only read as an input by group k/ds <= k, already summed */
		work[2 * k] = wrap16(si);
		work[2 * k + 1] = wrap16(sq);
	}
	for (k = outs; k < pairs; k++) {
		work[2 * k] = 0;
		work[2 * k + 1] = 0;
	}
}

void oracle_scan_read(const oracle_cfg_t *cfg, const uint8_t *buf8, int16_t *work,
		      int64_t *avg, int *samples)
{
	int n = 1 << cfg->bin_e, ds = cfg->downsample, ds_p = cfg->downsample_passes;
	int buf_len = cfg->buf_len, len, offset, j;

	if (n == 1) { /* :661-664 */
		avg[0] = oracle_rms_power(buf8, buf_len, avg[0], cfg->peak_hold);
		*samples += 1;
		return;
	}
	for (j = 0; j < buf_len; j++) /* :666-668 */
		work[j] = (int16_t)((int)buf8[j] - 127);
	if (cfg->boxcar && ds > 1) {
		/* boxcar_decimate reads inputs of group k before writing slot k and
		 * slot k < k*ds for k >= 1, so in-place is safe */
		boxcar_decimate(work, buf_len, ds);
	} else if (ds_p) { /* :683-691 */
		for (j = 0; j < ds_p; j++) {
			oracle_fifth_order(work, buf_len >> j);
			oracle_fifth_order(work + 1, (buf_len >> j) - 1);
		}
		if (cfg->comp_fir_size == 9 && ds_p <= 10) {
			oracle_generic_fir(work, buf_len >> ds_p, oracle_cic9(ds_p));
			oracle_generic_fir(work + 1, (buf_len >> ds_p) - 1, oracle_cic9(ds_p));
		}
	}
	len = buf_len / ds;
	oracle_remove_dc(work, len);          /* :692 */
	oracle_remove_dc(work + 1, len - 1);  /* :693 */
	for (offset = 0; offset < len; offset += 2 * n) { /* :695-718 */
		int16_t *blk = work + offset;
		for (j = 0; j < n; j++) {
			blk[2 * j] = wrap16((int32_t)blk[2 * j] * cfg->window[j]);
			blk[2 * j + 1] = wrap16((int32_t)blk[2 * j + 1] * cfg->window[j]);
		}
		oracle_fix_fft(blk, cfg->bin_e, cfg->sine, cfg->bin_e);
		for (j = 0; j < n; j++) {
			int64_t re = blk[2 * j], im = blk[2 * j + 1];
			int64_t p = re * re + im * im;
			if (!cfg->peak_hold)
				avg[j] += p;
			else if (p > avg[j])
				avg[j] = p;
		}
		*samples += ds;
	}
}

/* ---- report ----------------------------------------------------------- */

/* rtl_power.c:722-760 without the text formatting */
int oracle_epilogue(int64_t *avg, int bin_e, double crop, int rate, int samples, double *db)
{
	int len = 1 << bin_e, i, i1, i2, k = 0;
	double v;
	if (bin_e > 0) {
		avg[0] = avg[1];
		for (i = 0; i < len / 2; i++) {
			int64_t t = avg[i];
			avg[i] = avg[i + len / 2];
			avg[i + len / 2] = t;
		}
	}
	i1 = 0 + (int)((double)len * crop * 0.5);
	i2 = (len - 1) - (int)((double)len * crop * 0.5);
	for (i = i1; i <= i2; i++) {
		v = (double)avg[i];
		v /= (double)rate;
		v /= (double)samples;
		db[k++] = 10 * log10(v);
	}
	v = (double)avg[i2] / ((double)rate * (double)samples);
	if (bin_e == 0)
		v = (double)avg[0] / ((double)rate * (double)samples);
	db[k++] = 10 * log10(v);
	return k;
}

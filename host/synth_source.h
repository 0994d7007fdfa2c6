/*
 * synth_source -- deterministic synthetic / replay sample source that stands in
 * for an RTL2832 dongle (there is none on the GPU box).
 *
 * It implements the librtlsdr *streaming* signatures the scan path touches
 *   rtlsdr_read_sync      (reference include/rtl-sdr.h:470, src/librtlsdr.c:2689-2695)
 *   rtlsdr_read_async     (include/rtl-sdr.h:488-492,      src/librtlsdr.c:2826-2929)
 *   rtlsdr_cancel_async   (include/rtl-sdr.h:500,          src/librtlsdr.c:2932-2952)
 *   rtlsdr_set/get_center_freq, rtlsdr_reset_buffer, rtlsdr_set_sample_rate
 * plus the handful of housekeeping calls rtl_power/convenience link against,
 * so that the unmodified reference object and the GPU host program see the
 * same bytes.
 *
 * Bytes are a pure function of (mode, seed, hop, pass, byte offset):
 *   hop  = index of the current centre frequency in the registered hop table
 *   pass = number of data reads already served for that hop
 * The first read after every rtlsdr_set_center_freq() is the tuner-settling
 * dump the reference throws away (rtl_power.c:542-552); it returns 0x7F bytes
 * and does not advance the pass counter.
 */
#ifndef SYNTH_SOURCE_H
#define SYNTH_SOURCE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtlsdr_dev rtlsdr_dev_t;
typedef void (*rtlsdr_read_async_cb_t)(unsigned char *buf, uint32_t len, void *ctx);

enum synth_mode {
	SYNTH_XORSHIFT = 0,  /* uniform PRNG bytes, SURVEY.md 8(c) generator */
	SYNTH_COUNTER  = 1,  /* RTL2832 test mode: b[i] = i & 0xFF (rtl_test.c:119-141) */
	SYNTH_CONST    = 2,  /* every byte = param (127 = muted, 0 / 255 = saturation) */
	SYNTH_BIASED   = 3,  /* clamp(PRNG + param): exercises the remove_dc subtract branch */
	SYNTH_TONE     = 4,  /* clipped complex tone, param = amplitude, period from seed */
	SYNTH_REPLAY   = 5   /* bytes taken from a caller-supplied pool, read r -> pool[r % n] */
};

/* Configure the source behind `dev` (NULL = the process-wide default device). */
void synth_configure(rtlsdr_dev_t *dev, int mode, uint64_t seed, int param);
/* Register the hop table (centre frequencies in scan order). */
void synth_set_hops(rtlsdr_dev_t *dev, const int *freqs, int count);
/* Bytes per generated block (= one pass of one hop); rtl_power reads exactly
 * buf_len bytes per hop visit, so set this to buf_len.  Default 16384. */
void synth_set_block_len(rtlsdr_dev_t *dev, size_t block_len);
/* Replay pool: n_reads blocks of read_len bytes; not copied, caller keeps it alive. */
void synth_set_replay(rtlsdr_dev_t *dev, const uint8_t *pool, size_t read_len, size_t n_reads);
/* Restart all per-hop pass counters (and forget the tuned frequency). */
void synth_rewind(rtlsdr_dev_t *dev);
/* Fill `out` with the bytes read (hop, pass) would return. Pure function. */
void synth_fill(rtlsdr_dev_t *dev, int hop, uint64_t pass, uint8_t *out, size_t len);
/* Stateless variant used by tests and by the GPU bench to build device input. */
void synth_generate(int mode, uint64_t seed, int param, int tune_count,
		    int hop, uint64_t pass, uint8_t *out, size_t len);
/* Number of data reads served so far (all hops). */
uint64_t synth_reads_served(rtlsdr_dev_t *dev);

/* librtlsdr-compatible entry points (same names and argument meaning). */
uint32_t rtlsdr_get_device_count(void);
const char *rtlsdr_get_device_name(uint32_t index);
int rtlsdr_get_device_usb_strings(uint32_t index, char *manufact, char *product, char *serial);
int rtlsdr_open(rtlsdr_dev_t **dev, uint32_t index);
int rtlsdr_close(rtlsdr_dev_t *dev);
int rtlsdr_set_center_freq(rtlsdr_dev_t *dev, uint32_t freq);
uint32_t rtlsdr_get_center_freq(rtlsdr_dev_t *dev);
int rtlsdr_set_sample_rate(rtlsdr_dev_t *dev, uint32_t rate);
uint32_t rtlsdr_get_sample_rate(rtlsdr_dev_t *dev);
int rtlsdr_reset_buffer(rtlsdr_dev_t *dev);
int rtlsdr_read_sync(rtlsdr_dev_t *dev, void *buf, int len, int *n_read);
int rtlsdr_read_async(rtlsdr_dev_t *dev, rtlsdr_read_async_cb_t cb, void *ctx,
		      uint32_t buf_num, uint32_t buf_len);
int rtlsdr_cancel_async(rtlsdr_dev_t *dev);

#ifdef __cplusplus
}
#endif
#endif

/*
 * synth_source.c -- deterministic synthetic / replay stand-in for librtlsdr's
 * streaming half.  See synth_source.h for the contract.
 *
 * Semantics kept from the reference (file:line in /root/reference):
 *  - read_sync: one blocking read of `len` bytes, returns 0, *n_read = len
 *    (src/librtlsdr.c:2689-2695); the caller owns the buffer.
 *  - read_async: buf_num (default 15) buffers of buf_len bytes (default 32768,
 *    must be a multiple of 512 or the default is used), callback on the
 *    CALLING thread, the same buffer is re-armed as soon as the callback
 *    returns, -1 on NULL device, -2 if already streaming
 *    (src/librtlsdr.c:2826-2929, defaults :407-408).
 *  - cancel_async: 0 when streaming, -2 otherwise (src/librtlsdr.c:2932-2952).
 */
#include "synth_source.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SYNTH_MAX_HOPS      3000   /* = MAX_TUNES, rtl_power.c:113 */
#define ASYNC_DEFAULT_NUM   15     /* librtlsdr.c:407 */
#define ASYNC_DEFAULT_LEN   (64 * 512) /* librtlsdr.c:408 */
#define SETTLE_BYTE         0x7F

enum async_state { ASYNC_INACTIVE = 0, ASYNC_CANCELING, ASYNC_RUNNING };

struct rtlsdr_dev {
	int mode;
	uint64_t seed;
	int param;
	int hop_freq[SYNTH_MAX_HOPS];
	int hop_count;
	uint64_t hop_offset[SYNTH_MAX_HOPS]; /* bytes of the hop's stream served */
	size_t block_len;                    /* bytes per generated block (= one pass) */
	const uint8_t *pool;
	size_t pool_read_len, pool_reads;
	uint32_t freq, rate;
	int cur_hop;
	int settle_pending;
	uint64_t reads;
	volatile int async_status;
};

static struct rtlsdr_dev g_default_dev = { .block_len = 16384, .cur_hop = 0 };

static struct rtlsdr_dev *resolve(rtlsdr_dev_t *dev)
{
	return dev ? dev : &g_default_dev;
}

static inline uint64_t xs_next(uint64_t s)
{
	s ^= s << 13;
	s ^= s >> 7;
	s ^= s << 17;
	return s;
}

static uint64_t xs_seed(uint64_t seed, uint64_t r)
{
	/* s = 0x9E3779B97F4A7C15 ^ (r * 0x100000001B3); `seed`
	 * is folded in so seed 0 reproduces the survey's known-answer rows. */
	uint64_t s = (0x9E3779B97F4A7C15ULL + seed * 0xD1B54A32D192ED03ULL)
		     ^ (r * 0x100000001B3ULL);
	if (s == 0)
		s = 0x2545F4914F6CDD1DULL; /* xorshift must not start at 0 */
	return s;
}

void synth_generate(int mode, uint64_t seed, int param, int tune_count,
		    int hop, uint64_t pass, uint8_t *out, size_t len)
{
	size_t i;
	uint64_t r = pass * (uint64_t)(tune_count > 0 ? tune_count : 1) + (uint64_t)hop;
	uint64_t s;
	switch (mode) {
	case SYNTH_COUNTER:
		for (i = 0; i < len; i++)
			out[i] = (uint8_t)(i & 0xFF);
		break;
	case SYNTH_CONST:
		memset(out, param & 0xFF, len);
		break;
	case SYNTH_BIASED:
		s = xs_seed(seed, r);
		for (i = 0; i < len; i++) {
			int v;
			s = xs_next(s);
			v = (int)((s >> 32) & 0xFF) + param;
			out[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
		}
		break;
	case SYNTH_TONE: {
		/* clipped complex exponential; period and phase vary with (seed, r)
		 * so that different hops/passes land on different bins */
		double amp = (double)param;
		double cyc = 3.0 + (double)((seed * 7 + r * 13) % 509) + 0.125 * (double)(r % 8);
		double ph0 = 0.785398163397448309616 * (double)(r % 8);
		for (i = 0; i + 1 < len; i += 2) {
			double ph = ph0 + 6.283185307179586476925 * (double)(i / 2) / cyc;
			double vi = 127.0 + amp * cos(ph), vq = 127.0 + amp * sin(ph);
			long li = lround(vi), lq = lround(vq);
			out[i]     = (uint8_t)(li < 0 ? 0 : (li > 255 ? 255 : li));
			out[i + 1] = (uint8_t)(lq < 0 ? 0 : (lq > 255 ? 255 : lq));
		}
		if (len & 1)
			out[len - 1] = 127;
		break;
	}
	case SYNTH_XORSHIFT:
	default:
		s = xs_seed(seed, r);
		for (i = 0; i < len; i++) {
			s = xs_next(s);
			out[i] = (uint8_t)((s >> 32) & 0xFF);
		}
		break;
	}
}

/* ---- bulk generation (bench / sweep drivers fill whole sweep cubes) -------- */

#include <pthread.h>

struct cube_job {
	int mode, param, tune_count, hop_first, hop_count, worker, workers;
	uint64_t seed, pass_first, passes;
	uint8_t *out;
	size_t pass_stride, hop_stride, len;
};

static void *cube_worker(void *arg)
{
	const struct cube_job *j = (const struct cube_job *)arg;
	const uint64_t total = j->passes * (uint64_t)j->hop_count;
	uint64_t i;
	for (i = (uint64_t)j->worker; i < total; i += (uint64_t)j->workers) {
		const uint64_t p = i / (uint64_t)j->hop_count;
		const int k = (int)(i % (uint64_t)j->hop_count);
		synth_generate(j->mode, j->seed, j->param, j->tune_count, j->hop_first + k, j->pass_first + p,
			       j->out + (size_t)p * j->pass_stride + (size_t)k * j->hop_stride, j->len);
	}
	return NULL;
}

void synth_generate_cube(int mode, uint64_t seed, int param, int tune_count, int hop_first, int hop_count,
			 uint64_t pass_first, uint64_t passes, uint8_t *out, size_t pass_stride,
			 size_t hop_stride, size_t len, int threads)
{
	struct cube_job jobs[64];
	pthread_t th[64];
	int i, started = 0;
	if (threads < 1)
		threads = 1;
	if (threads > 64)
		threads = 64;
	for (i = 0; i < threads; i++) {
		struct cube_job j = { mode, param, tune_count, hop_first, hop_count, i, threads,
				      seed, pass_first, passes, out, pass_stride, hop_stride, len };
		jobs[i] = j;
	}
	for (i = 1; i < threads; i++) {
		if (pthread_create(&th[i], NULL, cube_worker, &jobs[i]) != 0)
			break;
		started = i;
	}
	/* workers that could not be started: their share is done here */
	for (i = started + 1; i < threads; i++)
		cube_worker(&jobs[i]);
	cube_worker(&jobs[0]);
	for (i = 1; i <= started; i++)
		pthread_join(th[i], NULL);
}

uint64_t synth_fnv1a_int64(const int64_t *words, size_t count, uint64_t h)
{
	size_t i;
	for (i = 0; i < count; i++) {
		h ^= (uint64_t)words[i];
		h *= 1099511628211ULL;
	}
	return h;
}

void synth_configure(rtlsdr_dev_t *dev, int mode, uint64_t seed, int param)
{
	struct rtlsdr_dev *d = resolve(dev);
	d->mode = mode;
	d->seed = seed;
	d->param = param;
	synth_rewind(d);
}

void synth_set_hops(rtlsdr_dev_t *dev, const int *freqs, int count)
{
	struct rtlsdr_dev *d = resolve(dev);
	if (count > SYNTH_MAX_HOPS)
		count = SYNTH_MAX_HOPS;
	if (count < 0)
		count = 0;
	if (count)
		memcpy(d->hop_freq, freqs, (size_t)count * sizeof(int));
	d->hop_count = count;
	synth_rewind(d);
}

void synth_set_block_len(rtlsdr_dev_t *dev, size_t block_len)
{
	struct rtlsdr_dev *d = resolve(dev);
	d->block_len = block_len ? block_len : 16384;
}

void synth_set_replay(rtlsdr_dev_t *dev, const uint8_t *pool, size_t read_len, size_t n_reads)
{
	struct rtlsdr_dev *d = resolve(dev);
	d->pool = pool;
	d->pool_read_len = read_len;
	d->pool_reads = n_reads;
	d->mode = SYNTH_REPLAY;
	d->block_len = read_len ? read_len : d->block_len;
	synth_rewind(d);
}

void synth_rewind(rtlsdr_dev_t *dev)
{
	struct rtlsdr_dev *d = resolve(dev);
	memset(d->hop_offset, 0, sizeof(d->hop_offset));
	d->freq = 0;
	d->cur_hop = 0;
	d->settle_pending = 0;
	d->reads = 0;
}

uint64_t synth_reads_served(rtlsdr_dev_t *dev)
{
	return resolve(dev)->reads;
}

static void fill_block(struct rtlsdr_dev *d, int hop, uint64_t block, uint8_t *out, size_t len)
{
	int tc = d->hop_count > 0 ? d->hop_count : 1;
	if (d->mode == SYNTH_REPLAY) {
		if (d->pool && d->pool_reads) {
			uint64_t r = (block * (uint64_t)tc + (uint64_t)hop) % d->pool_reads;
			size_t n = len < d->pool_read_len ? len : d->pool_read_len;
			memcpy(out, d->pool + r * d->pool_read_len, n);
			if (n < len)
				memset(out + n, SETTLE_BYTE, len - n);
		} else {
			memset(out, SETTLE_BYTE, len);
		}
		return;
	}
	synth_generate(d->mode, d->seed, d->param, tc, hop, block, out, len);
}

void synth_fill(rtlsdr_dev_t *dev, int hop, uint64_t pass, uint8_t *out, size_t len)
{
	fill_block(resolve(dev), hop, pass, out, len);
}

/* Serve the next `len` bytes of the current hop's stream.  The stream is the
 * concatenation of block_len-byte blocks, block b being pass b of that hop. */
static void serve(struct rtlsdr_dev *d, uint8_t *out, size_t len)
{
	int hop = d->cur_hop;
	uint64_t off = d->hop_offset[hop];
	size_t bl = d->block_len ? d->block_len : 16384;
	if (off % bl == 0 && len == bl) {
		fill_block(d, hop, off / bl, out, len); /* the rtl_power case */
	} else {
		uint8_t *tmp = (uint8_t *)malloc(bl);
		size_t done = 0;
		while (tmp && done < len) {
			uint64_t b = (off + done) / bl;
			size_t o = (size_t)((off + done) % bl);
			size_t n = bl - o;
			if (n > len - done)
				n = len - done;
			fill_block(d, hop, b, tmp, bl);
			memcpy(out + done, tmp + o, n);
			done += n;
		}
		free(tmp);
	}
	d->hop_offset[hop] = off + len;
	d->reads++;
}

/* ---- librtlsdr-compatible surface ------------------------------------- */

uint32_t rtlsdr_get_device_count(void) { return 1; }

const char *rtlsdr_get_device_name(uint32_t index)
{
	return index == 0 ? "Synthetic RTL2832U replay source" : "";
}

int rtlsdr_get_device_usb_strings(uint32_t index, char *manufact, char *product, char *serial)
{
	if (index != 0)
		return -1;
	if (manufact) strcpy(manufact, "synthetic");
	if (product) strcpy(product, "replay");
	if (serial) strcpy(serial, "00000001");
	return 0;
}

int rtlsdr_open(rtlsdr_dev_t **dev, uint32_t index)
{
	if (!dev || index != 0)
		return -1;
	*dev = &g_default_dev;
	return 0;
}

int rtlsdr_close(rtlsdr_dev_t *dev)
{
	return dev ? 0 : -1;
}

int rtlsdr_set_center_freq(rtlsdr_dev_t *dev, uint32_t freq)
{
	struct rtlsdr_dev *d = resolve(dev);
	int i;
	d->freq = freq;
	d->cur_hop = 0;
	for (i = 0; i < d->hop_count; i++) {
		if ((uint32_t)d->hop_freq[i] == freq) {
			d->cur_hop = i;
			break;
		}
	}
	d->settle_pending = 1;
	return 0;
}

uint32_t rtlsdr_get_center_freq(rtlsdr_dev_t *dev)
{
	return resolve(dev)->freq;
}

int rtlsdr_set_sample_rate(rtlsdr_dev_t *dev, uint32_t rate)
{
	resolve(dev)->rate = rate;
	return 0;
}

uint32_t rtlsdr_get_sample_rate(rtlsdr_dev_t *dev)
{
	return resolve(dev)->rate;
}

int rtlsdr_reset_buffer(rtlsdr_dev_t *dev)
{
	(void)dev;
	return 0;
}

int rtlsdr_read_sync(rtlsdr_dev_t *dev, void *buf, int len, int *n_read)
{
	struct rtlsdr_dev *d = resolve(dev);
	if (!buf || len < 0)
		return -1;
	if (d->settle_pending) {
		/* what the tuner produces while the PLL settles: thrown away by
		 * retune() (rtl_power.c:548-551); not part of any hop's stream */
		d->settle_pending = 0;
		memset(buf, SETTLE_BYTE, (size_t)len);
	} else {
		serve(d, (uint8_t *)buf, (size_t)len);
	}
	if (n_read)
		*n_read = len;
	return 0;
}

int rtlsdr_read_async(rtlsdr_dev_t *dev, rtlsdr_read_async_cb_t cb, void *ctx,
		      uint32_t buf_num, uint32_t buf_len)
{
	struct rtlsdr_dev *d;
	uint8_t **ring;
	uint32_t i, n, len;
	if (!dev)
		return -1;
	d = dev;
	if (d->async_status != ASYNC_INACTIVE)
		return -2;
	d->async_status = ASYNC_RUNNING;
	n = buf_num > 0 ? buf_num : ASYNC_DEFAULT_NUM;
	len = (buf_len > 0 && buf_len % 512 == 0) ? buf_len : ASYNC_DEFAULT_LEN;
	ring = (uint8_t **)calloc(n, sizeof(*ring));
	if (!ring) {
		d->async_status = ASYNC_INACTIVE;
		return -1;
	}
	for (i = 0; i < n; i++) {
		ring[i] = (uint8_t *)malloc(len);
		if (!ring[i])
			d->async_status = ASYNC_CANCELING;
	}
	d->settle_pending = 0; /* a stream has no settle dump; the user skips samples */
	i = 0;
	while (d->async_status == ASYNC_RUNNING) {
		serve(d, ring[i], len);
		if (cb)
			cb(ring[i], len, ctx); /* buffer is re-armed (overwritten) afterwards */
		i = (i + 1) % n;
	}
	for (i = 0; i < n; i++)
		free(ring[i]);
	free(ring);
	d->async_status = ASYNC_INACTIVE;
	return 0;
}

int rtlsdr_cancel_async(rtlsdr_dev_t *dev)
{
	if (!dev)
		return -1;
	if (dev->async_status == ASYNC_RUNNING) {
		dev->async_status = ASYNC_CANCELING;
		return 0;
	}
	return -2;
}

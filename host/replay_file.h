/*
 * replay_file -- recorded IQ as a sample source.  Three on-disk
 * layouts are recognised, all carrying interleaved unsigned 8-bit I,Q:
 *   raw      what `rtl_sdr file` writes (reference src/rtl_sdr.c:97-121)
 *   wav      what `rtl_sdr -w` / wavewrite.c writes: RIFF/WAVE header, "data" chunk
 *            (reference src/convenience/wavewrite.c:120-246)
 *   rtl_tcp  a captured rtl_tcp stream: 12-byte dongle_info ("RTL0", tuner type,
 *            gain count, big endian) followed by raw IQ (protocol_rtl_tcp.txt:22-48)
 */
#ifndef REPLAY_FILE_H
#define REPLAY_FILE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum replay_format { REPLAY_RAW = 0, REPLAY_WAV = 1, REPLAY_RTL_TCP = 2 };

typedef struct replay_info {
	int format;
	uint64_t payload_offset;  /* first IQ byte */
	uint64_t payload_bytes;   /* IQ bytes available */
	uint32_t sample_rate;     /* WAV only, else 0 */
	uint32_t tuner_type;      /* rtl_tcp only */
	uint32_t gain_count;      /* rtl_tcp only */
} replay_info_t;

/* Inspect `path`.  Returns 0, -1 if the file cannot be read, -2 on a malformed header. */
int replay_probe(const char *path, replay_info_t *info);

/* Load the payload as whole reads of read_len bytes (a trailing partial read is dropped).
 * Returns a malloc'd buffer of *n_reads * read_len bytes, or NULL. */
uint8_t *replay_load(const char *path, size_t read_len, size_t *n_reads, replay_info_t *info);

#ifdef __cplusplus
}
#endif
#endif

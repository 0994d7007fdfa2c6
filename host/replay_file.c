#include "replay_file.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint32_t le32(const unsigned char *p)
{
	return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

static uint32_t be32(const unsigned char *p)
{
	return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

int replay_probe(const char *path, replay_info_t *info)
{
	unsigned char head[12];
	long size;
	FILE *f;
	if (!path || !info)
		return -1;
	memset(info, 0, sizeof(*info));
	f = fopen(path, "rb");
	if (!f)
		return -1;
	fseek(f, 0, SEEK_END);
	size = ftell(f);
	fseek(f, 0, SEEK_SET);
	if (size < 0 || fread(head, 1, sizeof(head), f) != sizeof(head)) {
		/* shorter than any header: raw */
		info->format = REPLAY_RAW;
		info->payload_bytes = size > 0 ? (uint64_t)size : 0;
		fclose(f);
		return 0;
	}
	if (memcmp(head, "RIFF", 4) == 0 && memcmp(head + 8, "WAVE", 4) == 0) {
		/* walk the chunks up to "data"; "fmt " gives the sample rate */
		long pos = 12;
		info->format = REPLAY_WAV;
		for (;;) {
			unsigned char ch[8];
			uint32_t len;
			if (fseek(f, pos, SEEK_SET) != 0 || fread(ch, 1, 8, f) != 8) {
				fclose(f);
				return -2;
			}
			len = le32(ch + 4);
			if (memcmp(ch, "fmt ", 4) == 0) {
				unsigned char fmt[16];
				if (len >= 16 && fread(fmt, 1, 16, f) == 16)
					info->sample_rate = le32(fmt + 4);
			} else if (memcmp(ch, "data", 4) == 0) {
				info->payload_offset = (uint64_t)pos + 8;
				/* a recorder that was interrupted leaves dataSize = 0 (wavewrite.c:222): use the file size */
				if (len == 0 || (uint64_t)pos + 8 + len > (uint64_t)size)
					info->payload_bytes = (uint64_t)size - info->payload_offset;
				else
					info->payload_bytes = len;
				fclose(f);
				return 0;
			}
			pos += 8 + (long)len + (long)(len & 1);
			if (pos >= size) {
				fclose(f);
				return -2;
			}
		}
	}
	if (memcmp(head, "RTL0", 4) == 0) {
		info->format = REPLAY_RTL_TCP;
		info->tuner_type = be32(head + 4);
		info->gain_count = be32(head + 8);
		info->payload_offset = 12;
		info->payload_bytes = (uint64_t)size - 12;
		fclose(f);
		return 0;
	}
	info->format = REPLAY_RAW;
	info->payload_bytes = (uint64_t)size;
	fclose(f);
	return 0;
}

uint8_t *replay_load(const char *path, size_t read_len, size_t *n_reads, replay_info_t *info)
{
	replay_info_t local;
	uint8_t *buf;
	size_t n;
	FILE *f;
	if (!info)
		info = &local;
	if (!n_reads || read_len == 0 || replay_probe(path, info) != 0)
		return NULL;
	n = (size_t)(info->payload_bytes / read_len);
	*n_reads = n;
	if (n == 0)
		return NULL;
	f = fopen(path, "rb");
	if (!f)
		return NULL;
	buf = (uint8_t *)malloc(n * read_len);
	if (buf && (fseek(f, (long)info->payload_offset, SEEK_SET) != 0 || fread(buf, read_len, n, f) != n)) {
		free(buf);
		buf = NULL;
	}
	fclose(f);
	return buf;
}

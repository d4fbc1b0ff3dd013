/*
 * snappy_oracle.c -- CPU restatement of the csnappy hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * codec in csnappy_b200/csrc.  Nothing under csnappy_b200/ links, loads or
 * calls it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may.  It is written from the behavioural description
 * of the reference, index based and byte oriented,
 * not from the reference's pointer-walking code.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   - tests/golden/ fixtures (urls.10K -> urls.10K.snappy at wm 15, the
 *     per-wm size table, baddata3 => -5, unaligned_uint64 decode, the
 *     appendix-B error matrix), and
 *   - the unmodified reference compiled into oracle/_ref/libcsnappy_ref.so
 *     (oracle/Makefile) on random / periodic / edge-size inputs.
 *
 * Reference anchors (file:line under /root/reference):
 *   oracle_max_compressed_length     csnappy_compress.c:612-616
 *   oracle_compress_fragment         csnappy_compress.c:469-606
 *     hash                           csnappy_compress.c:228-236
 *     match extension                csnappy_compress.c:252-295
 *     literal emission               csnappy_compress.c:332-371
 *     copy emission / split rule     csnappy_compress.c:373-415
 *   oracle_compress                  csnappy_compress.c:46-73, 621-656
 *   oracle_get_uncompressed_length   csnappy_decompress.c:45-71
 *   oracle_decompress_noheader       csnappy_decompress.c:319-387
 *     copy validity / space order    csnappy_decompress.c:295-317
 *     literal checks                 csnappy_decompress.c:358-380
 *   oracle_decompress                csnappy_decompress.c:394-411
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define E_OK 0
#define E_HEADER_BAD (-1)
#define E_OUTPUT_INSUF (-2)
#define E_OUTPUT_OVERRUN (-3)
#define E_DATA_MALFORMED (-5)

#define FRAGMENT_MAX 32768u
#define TAIL_MARGIN 15u

static uint32_t le32_at(const uint8_t *buf, uint32_t pos)
{
	return (uint32_t)buf[pos] | ((uint32_t)buf[pos + 1] << 8) |
	       ((uint32_t)buf[pos + 2] << 16) | ((uint32_t)buf[pos + 3] << 24);
}

uint32_t oracle_max_compressed_length(uint32_t n)
{
	return 32u + n + n / 6u;
}

/* literal [src, src+len) -> tag (+ length bytes) + payload; len >= 1 */
static uint32_t put_literal(uint8_t *out, uint32_t op, const uint8_t *in,
			    uint32_t src, uint32_t len)
{
	uint32_t v = len - 1;
	if (v < 60) {
		out[op++] = (uint8_t)(v << 2);
	} else {
		uint32_t tagpos = op++, nbytes = 0;
		while (v) {
			out[op++] = (uint8_t)(v & 0xff);
			v >>= 8;
			nbytes++;
		}
		out[tagpos] = (uint8_t)((59 + nbytes) << 2);
	}
	memcpy(out + op, in + src, len);
	return op + len;
}

/* one copy element, 4 <= len <= 64, offset < 65536 */
static uint32_t put_copy_piece(uint8_t *out, uint32_t op, uint32_t offset,
			       uint32_t len)
{
	if (len < 12 && offset < 2048) {
		out[op++] = (uint8_t)(1 | ((len - 4) << 2) | ((offset >> 8) << 5));
		out[op++] = (uint8_t)(offset & 0xff);
	} else {
		out[op++] = (uint8_t)(2 | ((len - 1) << 2));
		out[op++] = (uint8_t)(offset & 0xff);
		out[op++] = (uint8_t)(offset >> 8);
	}
	return op;
}

static uint32_t put_copy(uint8_t *out, uint32_t op, uint32_t offset,
			 uint32_t len)
{
	while (len >= 68) {
		op = put_copy_piece(out, op, offset, 64);
		len -= 64;
	}
	if (len > 64) {
		op = put_copy_piece(out, op, offset, 60);
		len -= 60;
	}
	return put_copy_piece(out, op, offset, len);
}

/*
 * Greedy single-probe parse of one fragment (no length header).
 * Returns the number of bytes written to out.  wm = log2(table bytes).
 */
uint32_t oracle_compress_fragment(const uint8_t *in, uint32_t n, uint8_t *out,
				  int wm)
{
	uint16_t table[1u << 15];
	const int shift = 33 - wm;
	uint32_t op = 0, next_emit = 0, ip, ip_limit;

	if (n < TAIL_MARGIN) {
		if (n)
			op = put_literal(out, op, in, 0, n);
		return op;
	}
	memset(table, 0, (size_t)1 << wm);
	ip_limit = n - TAIL_MARGIN;
	ip = 1;

	for (;;) {
		/* scan: probe stride grows by one every 32 probes */
		uint32_t probes = 32, cand, hit = 0;
		for (;;) {
			uint32_t step = probes >> 5, h;
			probes++;
			if (ip + step > ip_limit)
				break;
			h = (le32_at(in, ip) * 0x1e35a7bdu) >> shift;
			cand = table[h];
			table[h] = (uint16_t)ip;
			if (le32_at(in, ip) == le32_at(in, cand)) {
				hit = 1;
				break;
			}
			ip += step;
		}
		if (!hit)
			break;

		op = put_literal(out, op, in, next_emit, ip - next_emit);
		for (;;) {
			uint32_t m = 4, h;
			while (ip + m < n && in[cand + m] == in[ip + m])
				m++;
			op = put_copy(out, op, ip - cand, m);
			ip += m;
			next_emit = ip;
			if (ip >= ip_limit)
				goto tail;
			h = (le32_at(in, ip - 1) * 0x1e35a7bdu) >> shift;
			table[h] = (uint16_t)(ip - 1);
			h = (le32_at(in, ip) * 0x1e35a7bdu) >> shift;
			cand = table[h];
			table[h] = (uint16_t)ip;
			if (le32_at(in, ip) != le32_at(in, cand))
				break;
		}
		ip++;
	}
tail:
	if (next_emit < n)
		op = put_literal(out, op, in, next_emit, n - next_emit);
	return op;
}

static uint32_t put_varint32(uint8_t *out, uint32_t v)
{
	uint32_t k = 0;
	while (v >= 128) {
		out[k++] = (uint8_t)(v | 0x80);
		v >>= 7;
	}
	out[k++] = (uint8_t)v;
	return k;
}

/* table-size rule for a short final chunk, csnappy_compress.c:638-646 */
int oracle_chunk_wm(uint32_t chunk_len, int wm)
{
	int ws;
	if (chunk_len >= FRAGMENT_MAX)
		return wm;
	for (ws = 9; ws < wm; ws++)
		if ((1u << (ws - 1)) >= chunk_len)
			break;
	return ws;
}

void oracle_compress(const uint8_t *in, uint32_t n, uint8_t *out,
		     uint32_t *out_len, int wm)
{
	uint32_t op = put_varint32(out, n), pos = 0;
	while (pos < n) {
		uint32_t chunk = n - pos < FRAGMENT_MAX ? n - pos : FRAGMENT_MAX;
		op += oracle_compress_fragment(in + pos, chunk, out + op,
					       oracle_chunk_wm(chunk, wm));
		pos += chunk;
	}
	*out_len = op;
}

int oracle_get_uncompressed_length(const uint8_t *src, uint32_t n,
				   uint32_t *result)
{
	uint32_t shift = 0, used = 0;
	*result = 0;
	for (;;) {
		uint8_t c;
		if (shift >= 32 || used == n)
			return E_HEADER_BAD;
		c = src[used++];
		*result |= (uint32_t)(c & 0x7f) << shift;
		if (c < 128)
			return (int)used;
		shift += 7;
	}
}

/*
 * Tag interpreter.  cap = *dst_len on entry; *dst_len is written only on
 * success.  A tag whose trailing bytes are cut off by end of input is -5
 * (the reference's x86 path is formally undefined there).
 */
int oracle_decompress_noheader(const uint8_t *src, uint32_t src_len,
			       uint8_t *dst, uint32_t *dst_len)
{
	const uint32_t cap = *dst_len;
	uint32_t pos = 0, produced = 0;

	while (pos < src_len) {
		uint32_t tag = src[pos++], kind = tag & 3, len, i;
		if (kind == 0) {
			len = (tag >> 2) + 1;
			if (len > 60) {
				uint32_t nb = len - 60, v = 0;
				if (src_len - pos < nb)
					return E_DATA_MALFORMED;
				for (i = 0; i < nb; i++)
					v |= (uint32_t)src[pos + i] << (8 * i);
				pos += nb;
				len = v + 1; /* wraps to 0 for 0xffffffff */
			}
			if ((int32_t)len >= 0) {
				if (src_len - pos < len)
					return E_DATA_MALFORMED;
			} else if (cap - produced >= len) {
				return E_DATA_MALFORMED; /* unreachable in practice */
			}
			if (cap - produced < len)
				return E_OUTPUT_OVERRUN;
			memcpy(dst + produced, src + pos, len);
			pos += len;
			produced += len;
		} else {
			uint32_t nb = kind == 1 ? 1 : (kind == 2 ? 2 : 4), off = 0;
			if (src_len - pos < nb)
				return E_DATA_MALFORMED;
			for (i = 0; i < nb; i++)
				off |= (uint32_t)src[pos + i] << (8 * i);
			pos += nb;
			if (kind == 1) {
				len = ((tag >> 2) & 7) + 4;
				off |= (tag >> 5) << 8;
			} else {
				len = (tag >> 2) + 1;
			}
			if (off == 0 || off > produced)
				return E_DATA_MALFORMED;
			if (cap - produced < len)
				return E_OUTPUT_OVERRUN;
			for (i = 0; i < len; i++)
				dst[produced + i] = dst[produced + i - off];
			produced += len;
		}
	}
	*dst_len = produced;
	return E_OK;
}

int oracle_decompress(const uint8_t *src, uint32_t src_len, uint8_t *dst,
		      uint32_t dst_len)
{
	uint32_t olen = 0;
	int n = oracle_get_uncompressed_length(src, src_len, &olen);
	if (n < 0)
		return n;
	if (olen > dst_len)
		return E_OUTPUT_INSUF;
	return oracle_decompress_noheader(src + n, src_len - (uint32_t)n, dst,
					  &olen);
}

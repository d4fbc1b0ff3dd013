/*
 * csnappy_shim.c -- plain-C host side of libcsnappy_b200.so.
 *
 * Exports (a) the six drop-in symbols of the reference's csnappy.h
 * (/root/reference/csnappy.h:30-119) on HOST pointers and (b) the batched entry
 * points of include/csnappy_batch.h.  This file contains no codec: it does
 * argument checks, the varint32 framing (csnappy_compress.c:46-73,
 * csnappy_decompress.c:45-71), buffer staging and CUDA launches.  Every byte
 * of compressed or decompressed payload is produced by the sm_100a kernels in
 * compress_kernel.cu / decompress_kernel.cu.  There is no CPU fallback: without
 * a usable device the decompress calls return CSNAPPY_E_DEVICE and the compress
 * calls (no error channel in the reference ABI) abort loudly.
 */
#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/csnappy.h"
#include "../../include/csnappy_batch.h"
#include "kernels.h"

#define SLOT_STRIDE_32K 38272u /* csnappy_max_compressed_length(32768) = 38261, rounded up to 16 */
#define HDR 16u		       /* device staging: [u32 out_len][i32 status][pad] then payload */
#define SMALL_CALL 65536u      /* single calls up to this size take the one-synchronisation path */

static __thread char tls_err[256];

static int set_err(const char *what, int cuda_err)
{
	snprintf(tls_err, sizeof(tls_err), "csnappy_b200: %s: %s", what,
		 cuda_err > 0 ? cudaGetErrorString((cudaError_t)cuda_err) : "invalid argument");
	return cuda_err > 0 ? CSNAPPY_E_DEVICE : CSNAPPY_E_BAD_ARG;
}

const char *csnappy_b200_last_error(void) { return tls_err; }

int csnappy_b200_device_ok(void)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		set_err("cudaGetDeviceCount", (int)e);
		cudaGetLastError();
		return 0;
	}
	return n > 0;
}

uint64_t csnappy_b200_kernel_launches(void) { return csb_launch_count(); }

/* ---- tuning knobs ------------------------------------------------------- */
static int g_compress_lanes, g_decompress_lanes, g_ctas_per_sm, g_stage_input, g_smem_kb;

int csnappy_b200_set_tuning(const char *key, int value)
{
	if (!key)
		return CSNAPPY_E_BAD_ARG;
	if (!strcmp(key, "compress_lanes") || !strcmp(key, "decompress_lanes")) {
		if (value != 0 && value != 8 && value != 16 && value != 32)
			return CSNAPPY_E_BAD_ARG;
		if (key[0] == 'c')
			g_compress_lanes = value;
		else
			g_decompress_lanes = value;
		return 0;
	}
	if (!strcmp(key, "decompress_stage_input")) {
		if (value < 0 || value > 4)
			return CSNAPPY_E_BAD_ARG;
		g_stage_input = value;
		return 0;
	}
	if (!strcmp(key, "decompress_smem_kb")) {
		if (value < 0 || value > 227)
			return CSNAPPY_E_BAD_ARG;
		g_smem_kb = value;
		return 0;
	}
	if (!strcmp(key, "ctas_per_sm")) {
		if (value < 0 || value > 8)
			return CSNAPPY_E_BAD_ARG;
		g_ctas_per_sm = value;
		return 0;
	}
	return CSNAPPY_E_BAD_ARG;
}

/* ---- pure host arithmetic ---------------------------------------------- */
uint32_t csnappy_max_compressed_length(uint32_t source_len) { return 32u + source_len + source_len / 6u; }

int csnappy_get_uncompressed_length(const char *start, uint32_t n, uint32_t *result)
{
	const uint8_t *p = (const uint8_t *)start;
	uint32_t used = 0, shift = 0;
	*result = 0;
	for (;;) {
		uint8_t c;
		if (shift >= 32 || used == n)
			return CSNAPPY_E_HEADER_BAD;
		c = p[used++];
		*result |= (uint32_t)(c & 0x7f) << shift;
		if (c < 128)
			return (int)used;
		shift += 7;
	}
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

/* ---- batched device-pointer entry points -------------------------------- */
int csnappy_batch_compress_fragments(const void *d_in, const uint64_t *d_in_off, uint64_t in_stride,
				     const uint32_t *d_in_len, uint32_t uniform_in_len, uint32_t n_blocks,
				     void *d_out, uint64_t out_stride, uint32_t *d_out_len,
				     int workmem_bytes_power_of_two, uint32_t flags, void *stream)
{
	struct csb_compress_args a;
	int e;
	if (workmem_bytes_power_of_two < 9 || workmem_bytes_power_of_two > 16)
		return set_err("workmem_bytes_power_of_two outside 9..16", 0);
	if (!d_in_len && uniform_in_len > CSB_FRAGMENT_MAX)
		return set_err("fragment longer than 32768 bytes", 0);
	if (n_blocks && (!d_in || !d_out || !d_out_len))
		return set_err("null buffer", 0);
	memset(&a, 0, sizeof(a));
	a.in = (const uint8_t *)d_in;
	a.in_off = d_in_off;
	a.in_stride = in_stride;
	a.in_len = d_in_len;
	a.uniform_len = uniform_in_len;
	a.n_blocks = n_blocks;
	a.out = (uint8_t *)d_out;
	a.out_stride = out_stride;
	a.out_len = d_out_len;
	a.wm = workmem_bytes_power_of_two;
	a.flags = flags;
	a.lanes = g_compress_lanes;
	a.ctas_per_sm = g_ctas_per_sm;
	e = csb_launch_compress(&a, (csb_stream_t)stream);
	return e ? set_err("compress launch", e) : 0;
}

int csnappy_batch_decompress(const void *d_in, const uint64_t *d_in_off, uint64_t in_stride,
			     const uint32_t *d_in_len, uint32_t n_blocks, void *d_out, uint64_t out_stride,
			     const uint32_t *d_out_cap, uint32_t uniform_out_cap, uint32_t *d_out_len,
			     int32_t *d_status, uint32_t flags, void *stream)
{
	struct csb_decompress_args a;
	int e;
	if (n_blocks && (!d_in || !d_in_len || !d_out_len || !d_status))
		return set_err("null buffer", 0);
	memset(&a, 0, sizeof(a));
	a.in = (const uint8_t *)d_in;
	a.in_off = d_in_off;
	a.in_stride = in_stride;
	a.in_len = d_in_len;
	a.n_blocks = n_blocks;
	a.out = (uint8_t *)d_out;
	a.out_stride = out_stride;
	a.out_cap = d_out_cap;
	a.uniform_cap = uniform_out_cap;
	a.out_len = d_out_len;
	a.status = d_status;
	a.flags = flags;
	a.lanes = g_decompress_lanes;
	a.stage_input = g_stage_input;
	a.smem_kb = g_smem_kb;
	a.ctas_per_sm = g_ctas_per_sm;
	e = csb_launch_decompress(&a, (csb_stream_t)stream);
	return e ? set_err("decompress launch", e) : 0;
}

int csnappy_batch_pack(const void *d_slots, uint64_t slot_stride, const uint32_t *d_len, uint32_t n_blocks,
		       void *d_packed, uint64_t *d_off, void *stream)
{
	int e;
	if (!d_off || (n_blocks && (!d_len || (d_packed && !d_slots))))
		return set_err("null buffer", 0);
	e = csb_launch_pack((const uint8_t *)d_slots, slot_stride, d_len, n_blocks, (uint8_t *)d_packed, d_off,
			    (csb_stream_t)stream);
	return e ? set_err("pack launch", e) : 0;
}

/* ---- staging context for the host-pointer calls ------------------------- */
struct buf {
	void *p;
	size_t cap;
};

static struct {
	pthread_mutex_t mu;
	int ready;
	cudaStream_t stream[3];
	struct buf d_in, d_out, d_aux, d_pack; /* device */
	struct buf h_pin;		       /* pinned host bounce buffer */
	struct buf d_in2[3], d_out2[3], d_aux2[3], d_ctr2[3], d_ctr;
} C = {.mu = PTHREAD_MUTEX_INITIALIZER};

static int ctx_init(void)
{
	int i, e;
	if (C.ready)
		return 0;
	for (i = 0; i < 3; i++)
		if ((e = (int)cudaStreamCreateWithFlags(&C.stream[i], cudaStreamNonBlocking)))
			return e;
	C.ready = 1;
	return 0;
}

static int grow_dev(struct buf *b, size_t need)
{
	int e;
	if (need <= b->cap)
		return 0;
	if (b->p && (e = (int)cudaFree(b->p)))
		return e;
	b->p = NULL;
	b->cap = 0;
	need = (need + (1u << 20)) & ~(size_t)((1u << 20) - 1);
	if ((e = (int)cudaMalloc(&b->p, need)))
		return e;
	b->cap = need;
	return 0;
}

static int grow_pin(struct buf *b, size_t need)
{
	int e;
	if (need <= b->cap)
		return 0;
	if (b->p && (e = (int)cudaFreeHost(b->p)))
		return e;
	b->p = NULL;
	b->cap = 0;
	need = (need + (1u << 16)) & ~(size_t)((1u << 16) - 1);
	if ((e = (int)cudaMallocHost(&b->p, need)))
		return e;
	b->cap = need;
	return 0;
}

#define TRY(what, expr)                          \
	do {                                     \
		int e__ = (int)(expr);           \
		if (e__) {                       \
			rc = set_err(what, e__); \
			goto out;                \
		}                                \
	} while (0)

static void die_no_error_channel(const char *fn)
{
	fprintf(stderr, "%s: %s (this entry point has no error return in the csnappy.h ABI; aborting)\n", fn, tls_err);
	abort();
}

/*
 * Core of both compress entry points.  framed = 0: one fragment, no header.
 * framed = 1: varint32 + 32 KiB fragments with the short-chunk table rule.
 * Returns bytes written to `out` (>= 0) or a negative error.
 */
static int64_t compress_host(const uint8_t *in, uint32_t n, uint8_t *out, int wm, int framed)
{
	int64_t rc = 0;
	uint32_t hdr = framed ? put_varint32(out, n) : 0;
	uint32_t n_frag = framed ? (uint32_t)(((uint64_t)n + CSB_FRAGMENT_MAX - 1) / CSB_FRAGMENT_MAX) : (n ? 1u : 0u);
	struct csb_compress_args a;
	cudaStream_t s;

	if (wm < 9 || wm > 16)
		return set_err("workmem_bytes_power_of_two outside 9..16", 0);
	if (!framed && n > CSB_FRAGMENT_MAX)
		return set_err("csnappy_compress_fragment: input longer than 32768 bytes", 0);
	if (n_frag == 0)
		return hdr;

	pthread_mutex_lock(&C.mu);
	TRY("stream create", ctx_init());
	s = C.stream[0];
	if (n_frag == 1) {
		/* one fragment (the per-page call of zram / block_compressor): pinned bounce buffers, the size travels
		 * in front of the slot, ONE stream synchronisation */
		const size_t in_pad = ((size_t)n + 63) & ~(size_t)63, slot = HDR + csnappy_max_compressed_length(n);
		uint8_t *pin;
		uint32_t clen;
		TRY("cudaMallocHost", grow_pin(&C.h_pin, in_pad + slot + 64));
		TRY("cudaMalloc(in)", grow_dev(&C.d_in, (size_t)n + 64));
		TRY("cudaMalloc(slots)", grow_dev(&C.d_out, SLOT_STRIDE_32K + HDR));
		TRY("cudaMalloc(ctr)", grow_dev(&C.d_ctr, 64));
		pin = (uint8_t *)C.h_pin.p;
		memcpy(pin, in, n);
		TRY("H2D", cudaMemcpyAsync(C.d_in.p, pin, n, cudaMemcpyHostToDevice, s));
		memset(&a, 0, sizeof(a));
		a.in = (const uint8_t *)C.d_in.p;
		a.in_stride = CSB_FRAGMENT_MAX;
		a.uniform_len = n;
		a.n_blocks = 1;
		a.out = (uint8_t *)C.d_out.p + HDR;
		a.out_stride = SLOT_STRIDE_32K;
		a.out_len = (uint32_t *)C.d_out.p;
		a.wm = wm;
		a.flags = framed ? CSNAPPY_BATCH_SHRINK_TABLE : 0;
		a.lanes = g_compress_lanes;
		a.ctas_per_sm = g_ctas_per_sm;
		a.counter = (uint32_t *)C.d_ctr.p;
		TRY("compress launch", csb_launch_compress(&a, (csb_stream_t)s));
		TRY("D2H", cudaMemcpyAsync(pin + in_pad, C.d_out.p, slot, cudaMemcpyDeviceToHost, s));
		TRY("sync", cudaStreamSynchronize(s));
		memcpy(&clen, pin + in_pad, 4);
		memcpy(out + hdr, pin + in_pad + HDR, clen);
		rc = (int64_t)hdr + clen;
		goto out;
	}
	TRY("cudaMalloc(in)", grow_dev(&C.d_in, (size_t)n + 64));
	TRY("cudaMalloc(slots)", grow_dev(&C.d_out, (size_t)n_frag * SLOT_STRIDE_32K));
	TRY("cudaMalloc(aux)", grow_dev(&C.d_aux, HDR + (size_t)n_frag * 4 + ((size_t)n_frag + 1) * 8 + 64));
	TRY("H2D", cudaMemcpyAsync(C.d_in.p, in, n, cudaMemcpyHostToDevice, s));

	memset(&a, 0, sizeof(a));
	a.in = (const uint8_t *)C.d_in.p;
	a.in_stride = CSB_FRAGMENT_MAX;
	a.uniform_len = n < CSB_FRAGMENT_MAX ? n : CSB_FRAGMENT_MAX;
	a.total_len = n;
	a.n_blocks = n_frag;
	a.out = (uint8_t *)C.d_out.p;
	a.out_stride = SLOT_STRIDE_32K;
	a.out_len = (uint32_t *)C.d_aux.p;
	a.wm = wm;
	a.flags = framed ? CSNAPPY_BATCH_SHRINK_TABLE : 0;
	a.lanes = g_compress_lanes;
	a.ctas_per_sm = g_ctas_per_sm;
	TRY("compress launch", csb_launch_compress(&a, (csb_stream_t)s));

	{
		uint64_t *d_off = (uint64_t *)((uint8_t *)C.d_aux.p + (((size_t)n_frag * 4 + 15) & ~(size_t)15));
		uint64_t total = 0;
		TRY("cudaMalloc(pack)", grow_dev(&C.d_pack, (size_t)n + (size_t)n / 6 + 32ull * n_frag + 64));
		TRY("pack launch", csb_launch_pack((const uint8_t *)C.d_out.p, SLOT_STRIDE_32K, (const uint32_t *)C.d_aux.p,
						   n_frag, (uint8_t *)C.d_pack.p, d_off, (csb_stream_t)s));
		TRY("D2H total", cudaMemcpyAsync(&total, d_off + n_frag, 8, cudaMemcpyDeviceToHost, s));
		TRY("sync", cudaStreamSynchronize(s));
		TRY("D2H data", cudaMemcpyAsync(out + hdr, C.d_pack.p, total, cudaMemcpyDeviceToHost, s));
		TRY("sync", cudaStreamSynchronize(s));
		rc = (int64_t)hdr + (int64_t)total;
	}
out:
	pthread_mutex_unlock(&C.mu);
	return rc;
}

char *csnappy_compress_fragment(const char *input, const uint32_t input_length, char *output,
				void *working_memory, const int workmem_bytes_power_of_two)
{
	int64_t r;
	(void)working_memory; /* the hash table lives in shared memory on the device */
	r = compress_host((const uint8_t *)input, input_length, (uint8_t *)output, workmem_bytes_power_of_two, 0);
	if (r < 0)
		die_no_error_channel("csnappy_compress_fragment");
	return output + r;
}

void csnappy_compress(const char *input, uint32_t input_length, char *compressed,
		      uint32_t *out_compressed_length, void *working_memory,
		      const int workmem_bytes_power_of_two)
{
	int64_t r;
	(void)working_memory;
	r = compress_host((const uint8_t *)input, input_length, (uint8_t *)compressed, workmem_bytes_power_of_two, 1);
	if (r < 0)
		die_no_error_channel("csnappy_compress");
	*out_compressed_length = (uint32_t)r;
}

/* raw decode of src[0..src_len) into dst with capacity cap; *produced set on success */
static int decompress_host(const uint8_t *src, uint32_t src_len, uint8_t *dst, uint32_t cap, uint32_t *produced)
{
	int rc = 0;
	struct csb_decompress_args a;
	cudaStream_t s;
	struct {
		uint32_t out_len;
		int32_t status;
	} res = {0, 0};
	uint32_t *d_res;

	pthread_mutex_lock(&C.mu);
	TRY("stream create", ctx_init());
	s = C.stream[0];
	if (src_len <= SMALL_CALL && cap <= SMALL_CALL) {
		/* one small block (the per-page call of zram / block_compressor): pinned bounce buffers, the input
		 * length travels in front of the input and [out_len, status] in front of the output, ONE synchronisation */
		const size_t in_bytes = HDR + src_len, in_pad = (in_bytes + 63) & ~(size_t)63, out_bytes = HDR + cap;
		uint8_t *pin;
		uint32_t *d_ihdr, *d_ohdr;
		TRY("cudaMallocHost", grow_pin(&C.h_pin, in_pad + out_bytes + 64));
		TRY("cudaMalloc(in)", grow_dev(&C.d_in, in_bytes + 64));
		TRY("cudaMalloc(out)", grow_dev(&C.d_out, out_bytes + 64));
		TRY("cudaMalloc(ctr)", grow_dev(&C.d_ctr, 64));
		pin = (uint8_t *)C.h_pin.p;
		memset(pin, 0, HDR);
		memcpy(pin, &src_len, 4);
		memcpy(pin + HDR, src, src_len);
		TRY("H2D", cudaMemcpyAsync(C.d_in.p, pin, in_bytes, cudaMemcpyHostToDevice, s));
		d_ihdr = (uint32_t *)C.d_in.p;
		d_ohdr = (uint32_t *)C.d_out.p;
		memset(&a, 0, sizeof(a));
		a.in = (const uint8_t *)C.d_in.p + HDR;
		a.in_len = d_ihdr;
		a.n_blocks = 1;
		a.out = (uint8_t *)C.d_out.p + HDR;
		a.uniform_cap = cap;
		a.out_len = d_ohdr;
		a.status = (int32_t *)(d_ohdr + 1);
		a.max_in_len = src_len;
		a.lanes = g_decompress_lanes;
		a.stage_input = g_stage_input;
		a.smem_kb = g_smem_kb;
		a.ctas_per_sm = g_ctas_per_sm;
		a.counter = (uint32_t *)C.d_ctr.p;
		TRY("decompress launch", csb_launch_decompress(&a, (csb_stream_t)s));
		TRY("D2H", cudaMemcpyAsync(pin + in_pad, C.d_out.p, out_bytes, cudaMemcpyDeviceToHost, s));
		TRY("sync", cudaStreamSynchronize(s));
		memcpy(&res, pin + in_pad, 8);
		rc = res.status;
		if (rc == 0) {
			memcpy(dst, pin + in_pad + HDR, res.out_len);
			*produced = res.out_len;
		}
		goto out;
	}
	TRY("cudaMalloc(in)", grow_dev(&C.d_in, (size_t)src_len + 64));
	TRY("cudaMalloc(out)", grow_dev(&C.d_out, (size_t)cap + 64));
	TRY("cudaMalloc(aux)", grow_dev(&C.d_aux, 64));
	d_res = (uint32_t *)C.d_aux.p;
	if (src_len)
		TRY("H2D", cudaMemcpyAsync(C.d_in.p, src, src_len, cudaMemcpyHostToDevice, s));
	TRY("H2D len", cudaMemcpyAsync(d_res + 2, &src_len, 4, cudaMemcpyHostToDevice, s));

	memset(&a, 0, sizeof(a));
	a.in = (const uint8_t *)C.d_in.p;
	a.in_stride = 0;
	a.in_len = d_res + 2;
	a.n_blocks = 1;
	a.out = (uint8_t *)C.d_out.p;
	a.out_stride = 0;
	a.uniform_cap = cap;
	a.out_len = d_res;
	a.status = (int32_t *)(d_res + 1);
	a.max_in_len = src_len;
	a.lanes = g_decompress_lanes;
	a.stage_input = g_stage_input;
	a.smem_kb = g_smem_kb;
	a.ctas_per_sm = g_ctas_per_sm;
	TRY("decompress launch", csb_launch_decompress(&a, (csb_stream_t)s));
	TRY("D2H result", cudaMemcpyAsync(&res, d_res, 8, cudaMemcpyDeviceToHost, s));
	TRY("sync", cudaStreamSynchronize(s));
	rc = res.status;
	if (rc == 0) {
		if (res.out_len)
			TRY("D2H data", cudaMemcpyAsync(dst, C.d_out.p, res.out_len, cudaMemcpyDeviceToHost, s));
		TRY("sync", cudaStreamSynchronize(s));
		*produced = res.out_len;
	}
out:
	pthread_mutex_unlock(&C.mu);
	return rc;
}

int csnappy_decompress_noheader(const char *src, uint32_t src_len, char *dst, uint32_t *dst_len)
{
	uint32_t produced = 0;
	int rc = decompress_host((const uint8_t *)src, src_len, (uint8_t *)dst, *dst_len, &produced);
	if (rc == 0)
		*dst_len = produced; /* written only on success, csnappy_decompress.c:385 */
	return rc;
}

int csnappy_decompress(const char *src, uint32_t src_len, char *dst, uint32_t dst_len)
{
	uint32_t olen = 0, produced = 0;
	int n = csnappy_get_uncompressed_length(src, src_len, &olen);
	if (n < 0)
		return n;
	if (olen > dst_len)
		return CSNAPPY_E_OUTPUT_INSUF;
	return decompress_host((const uint8_t *)src + n, src_len - (uint32_t)n, (uint8_t *)dst, olen, &produced);
}

/* ---- host-buffer batches: chunked, three streams, H2D / kernel / D2H overlap ---- */
#define NPIPE 3

static size_t pick_chunk_blocks(uint64_t in_stride, uint64_t out_stride, uint32_t n_blocks)
{
	/* ~32 MiB of the larger side per chunk keeps PCIe busy and the kernels full */
	uint64_t per = in_stride > out_stride ? in_stride : out_stride;
	uint64_t blocks = per ? (32ull << 20) / per : n_blocks;
	if (blocks < 1024)
		blocks = 1024;
	if (blocks > n_blocks)
		blocks = n_blocks;
	return (size_t)blocks;
}

int csnappy_batch_compress_fragments_host(const void *h_in, uint64_t in_stride, uint32_t uniform_in_len,
					  uint32_t n_blocks, void *h_out, uint64_t out_stride,
					  uint32_t *h_out_len, int workmem_bytes_power_of_two)
{
	int rc = 0, k;
	size_t chunk, done;
	if (workmem_bytes_power_of_two < 9 || workmem_bytes_power_of_two > 16)
		return set_err("workmem_bytes_power_of_two outside 9..16", 0);
	if (uniform_in_len > CSB_FRAGMENT_MAX || in_stride < uniform_in_len ||
	    out_stride < csnappy_max_compressed_length(uniform_in_len))
		return set_err("bad block geometry", 0);
	if (n_blocks == 0)
		return 0;
	if (!h_in || !h_out || !h_out_len)
		return set_err("null buffer", 0);

	pthread_mutex_lock(&C.mu);
	TRY("stream create", ctx_init());
	chunk = pick_chunk_blocks(in_stride, out_stride, n_blocks);
	for (k = 0; k < NPIPE; k++) {
		TRY("cudaMalloc(in)", grow_dev(&C.d_in2[k], chunk * in_stride + 64));
		TRY("cudaMalloc(out)", grow_dev(&C.d_out2[k], chunk * out_stride + 64));
		TRY("cudaMalloc(len)", grow_dev(&C.d_aux2[k], chunk * 4 + 64));
		TRY("cudaMalloc(ctr)", grow_dev(&C.d_ctr2[k], 64));
	}
	for (done = 0, k = 0; done < n_blocks; done += chunk, k = (k + 1) % NPIPE) {
		size_t nb = n_blocks - done < chunk ? n_blocks - done : chunk;
		cudaStream_t s = C.stream[k];
		struct csb_compress_args a;
		TRY("H2D", cudaMemcpyAsync(C.d_in2[k].p, (const uint8_t *)h_in + done * in_stride, nb * in_stride,
					   cudaMemcpyHostToDevice, s));
		memset(&a, 0, sizeof(a));
		a.in = (const uint8_t *)C.d_in2[k].p;
		a.in_stride = in_stride;
		a.uniform_len = uniform_in_len;
		a.n_blocks = (uint32_t)nb;
		a.out = (uint8_t *)C.d_out2[k].p;
		a.out_stride = out_stride;
		a.out_len = (uint32_t *)C.d_aux2[k].p;
		a.wm = workmem_bytes_power_of_two;
		a.lanes = g_compress_lanes;
		a.ctas_per_sm = g_ctas_per_sm;
		a.counter = (uint32_t *)C.d_ctr2[k].p;
		TRY("compress launch", csb_launch_compress(&a, (csb_stream_t)s));
		TRY("D2H data", cudaMemcpyAsync((uint8_t *)h_out + done * out_stride, C.d_out2[k].p, nb * out_stride,
						cudaMemcpyDeviceToHost, s));
		TRY("D2H len", cudaMemcpyAsync(h_out_len + done, C.d_aux2[k].p, nb * 4, cudaMemcpyDeviceToHost, s));
	}
	for (k = 0; k < NPIPE; k++)
		TRY("sync", cudaStreamSynchronize(C.stream[k]));
out:
	if (rc)
		for (k = 0; k < NPIPE; k++)
			if (C.ready)
				cudaStreamSynchronize(C.stream[k]);
	pthread_mutex_unlock(&C.mu);
	return rc;
}

int csnappy_batch_decompress_host(const void *h_in, uint64_t in_stride, const uint32_t *h_in_len,
				  uint32_t n_blocks, void *h_out, uint64_t out_stride, uint32_t uniform_out_cap,
				  uint32_t *h_out_len, int32_t *h_status, uint32_t flags)
{
	int rc = 0, k;
	size_t chunk, done;
	if (out_stride < uniform_out_cap)
		return set_err("bad block geometry", 0);
	if (n_blocks == 0)
		return 0;
	if (!h_in || !h_in_len || !h_out || !h_out_len || !h_status)
		return set_err("null buffer", 0);
	{
		/* a length beyond the stride would make the kernel read past the staged chunk (csnappy_bc_decompress_host
		 * checks its index the same way) */
		uint32_t i;
		for (i = 0; i < n_blocks; i++)
			if (h_in_len[i] > in_stride)
				return set_err("h_in_len[i] larger than in_stride", 0);
	}

	pthread_mutex_lock(&C.mu);
	TRY("stream create", ctx_init());
	chunk = pick_chunk_blocks(in_stride, out_stride, n_blocks);
	for (k = 0; k < NPIPE; k++) {
		TRY("cudaMalloc(in)", grow_dev(&C.d_in2[k], chunk * in_stride + 64));
		TRY("cudaMalloc(out)", grow_dev(&C.d_out2[k], chunk * out_stride + 64));
		TRY("cudaMalloc(aux)", grow_dev(&C.d_aux2[k], chunk * 12 + 64));
		TRY("cudaMalloc(ctr)", grow_dev(&C.d_ctr2[k], 64));
	}
	for (done = 0, k = 0; done < n_blocks; done += chunk, k = (k + 1) % NPIPE) {
		size_t nb = n_blocks - done < chunk ? n_blocks - done : chunk;
		cudaStream_t s = C.stream[k];
		uint32_t *d_ilen = (uint32_t *)C.d_aux2[k].p, *d_olen = d_ilen + chunk;
		int32_t *d_st = (int32_t *)(d_olen + chunk);
		struct csb_decompress_args a;
		TRY("H2D", cudaMemcpyAsync(C.d_in2[k].p, (const uint8_t *)h_in + done * in_stride, nb * in_stride,
					   cudaMemcpyHostToDevice, s));
		TRY("H2D len", cudaMemcpyAsync(d_ilen, h_in_len + done, nb * 4, cudaMemcpyHostToDevice, s));
		memset(&a, 0, sizeof(a));
		a.in = (const uint8_t *)C.d_in2[k].p;
		a.in_stride = in_stride;
		a.in_len = d_ilen;
		a.n_blocks = (uint32_t)nb;
		a.out = (uint8_t *)C.d_out2[k].p;
		a.out_stride = out_stride;
		a.uniform_cap = uniform_out_cap;
		a.out_len = d_olen;
		a.status = d_st;
		a.flags = flags;
		a.lanes = g_decompress_lanes;
		a.stage_input = g_stage_input;
		a.smem_kb = g_smem_kb;
		a.ctas_per_sm = g_ctas_per_sm;
		a.counter = (uint32_t *)C.d_ctr2[k].p;
		TRY("decompress launch", csb_launch_decompress(&a, (csb_stream_t)s));
		TRY("D2H data", cudaMemcpyAsync((uint8_t *)h_out + done * out_stride, C.d_out2[k].p, nb * out_stride,
						cudaMemcpyDeviceToHost, s));
		TRY("D2H len", cudaMemcpyAsync(h_out_len + done, d_olen, nb * 4, cudaMemcpyDeviceToHost, s));
		TRY("D2H status", cudaMemcpyAsync(h_status + done, d_st, nb * 4, cudaMemcpyDeviceToHost, s));
	}
	for (k = 0; k < NPIPE; k++)
		TRY("sync", cudaStreamSynchronize(C.stream[k]));
out:
	if (rc)
		for (k = 0; k < NPIPE; k++)
			if (C.ready)
				cudaStreamSynchronize(C.stream[k]);
	pthread_mutex_unlock(&C.mu);
	return rc;
}

/* ---- block_compressor-style container on host buffers ----------------------------------
 * Layout (reference block_compressor.c:275-394):
 *     [u32 nr_pages][u32 clen[nr_pages]][payload_0]...[payload_{nr_pages-1}]
 * payload_i is csnappy_compress_fragment(page_i, wm) or, when that is not smaller than the
 * page, the page itself with clen_i = its length (:316-318); the reader treats
 * clen_i == page_size as stored (:378).  Pages are compressed, size-scanned and packed on the
 * device chunk by chunk; four chunks are in flight so that H2D, kernels and D2H overlap, and
 * only COMPRESSED bytes cross the bus on the compressed side.
 */
#define BC_PIPE 4

struct bc_slot {
	cudaStream_t s;
	cudaEvent_t ev;
	struct buf d_in, d_slots, d_len, d_clen, d_off, d_packed, d_out, d_res, d_ctr;
	struct buf h_res; /* pinned: [u64 total] or [u32 out_len[n]][i32 status[n]] */
	int busy;
	uint64_t first;
	uint32_t n;
};
static struct bc_slot BC[BC_PIPE];
static int bc_ready;

static int bc_init(void)
{
	int i, e;
	if (bc_ready)
		return 0;
	for (i = 0; i < BC_PIPE; i++) {
		if ((e = (int)cudaStreamCreateWithFlags(&BC[i].s, cudaStreamNonBlocking)))
			return e;
		if ((e = (int)cudaEventCreateWithFlags(&BC[i].ev, cudaEventDisableTiming)))
			return e;
	}
	bc_ready = 1;
	return 0;
}

static uint32_t bc_chunk_pages(uint32_t page_size, uint64_t nr_pages)
{
	uint64_t pages = (32ull << 20) / page_size;
	if (pages < 256)
		pages = 256;
	if (pages > nr_pages)
		pages = nr_pages;
	return (uint32_t)pages;
}

uint64_t csnappy_bc_max_container_length(uint64_t input_length, uint32_t page_size)
{
	uint64_t nr = page_size ? (input_length + page_size - 1) / page_size : 0;
	return 4 + 4 * nr + input_length;
}

int csnappy_bc_compress_host(const void *h_in, uint64_t input_length, uint32_t page_size, void *h_container,
			     uint64_t container_capacity, uint64_t *container_length,
			     int workmem_bytes_power_of_two)
{
	int rc = 0, k;
	uint64_t nr, done, payload_pos, retire_next = 0, issued = 0;
	uint32_t chunk, out_stride;
	uint8_t *cont = (uint8_t *)h_container;
	if (workmem_bytes_power_of_two < 9 || workmem_bytes_power_of_two > 16)
		return set_err("workmem_bytes_power_of_two outside 9..16", 0);
	if (page_size == 0 || page_size > CSB_FRAGMENT_MAX)
		return set_err("page_size outside 1..32768", 0);
	nr = (input_length + page_size - 1) / page_size;
	if (nr > 0xffffffffull)
		return set_err("input too big", 0);
	if (!h_container || !container_length || (input_length && !h_in) ||
	    container_capacity < csnappy_bc_max_container_length(input_length, page_size))
		return set_err("container buffer missing or smaller than csnappy_bc_max_container_length", 0);
	{
		uint32_t nr32 = (uint32_t)nr;
		memcpy(cont, &nr32, 4);
	}
	payload_pos = 4 + 4 * nr;
	*container_length = payload_pos;
	if (nr == 0)
		return 0;
	chunk = bc_chunk_pages(page_size, nr);
	out_stride = (csnappy_max_compressed_length(page_size) + 15u) & ~15u;

	pthread_mutex_lock(&C.mu);
	TRY("stream create", bc_init());
	for (k = 0; k < BC_PIPE; k++) {
		struct bc_slot *b = &BC[k];
		TRY("cudaMalloc(in)", grow_dev(&b->d_in, (size_t)chunk * page_size + 64));
		TRY("cudaMalloc(slots)", grow_dev(&b->d_slots, (size_t)chunk * out_stride + 64));
		TRY("cudaMalloc(len)", grow_dev(&b->d_len, (size_t)chunk * 4 + 64));
		TRY("cudaMalloc(clen)", grow_dev(&b->d_clen, (size_t)chunk * 4 + 64));
		TRY("cudaMalloc(off)", grow_dev(&b->d_off, ((size_t)chunk + 1) * 8 + 64));
		TRY("cudaMalloc(packed)", grow_dev(&b->d_packed, (size_t)chunk * page_size + 64));
		TRY("cudaMallocHost(res)", grow_pin(&b->h_res, 64));
		TRY("cudaMalloc(ctr)", grow_dev(&b->d_ctr, 64));
		b->busy = 0;
	}
	for (done = 0; done < nr || retire_next < issued;) {
		/* retire the oldest chunk once two younger ones are queued (or nothing is left to queue) */
		if (retire_next < issued && (issued - retire_next > 2 || done >= nr)) {
			struct bc_slot *b = &BC[retire_next % BC_PIPE];
			uint64_t total;
			TRY("event sync", cudaEventSynchronize(b->ev));
			total = *(uint64_t *)b->h_res.p;
			TRY("D2H payload", cudaMemcpyAsync(cont + payload_pos, b->d_packed.p, total, cudaMemcpyDeviceToHost, b->s));
			payload_pos += total;
			retire_next++;
			continue;
		}
		{
			struct bc_slot *b = &BC[issued % BC_PIPE];
			uint64_t left_pages = nr - done, in_at = done * page_size;
			uint32_t nb = left_pages < chunk ? (uint32_t)left_pages : chunk;
			uint64_t in_bytes = input_length - in_at < (uint64_t)nb * page_size ? input_length - in_at
											      : (uint64_t)nb * page_size;
			struct csb_compress_args a;
			TRY("H2D", cudaMemcpyAsync(b->d_in.p, (const uint8_t *)h_in + in_at, in_bytes, cudaMemcpyHostToDevice, b->s));
			memset(&a, 0, sizeof(a));
			a.in = (const uint8_t *)b->d_in.p;
			a.in_stride = page_size;
			a.uniform_len = page_size;
			a.total_len = in_bytes;
			a.n_blocks = nb;
			a.out = (uint8_t *)b->d_slots.p;
			a.out_stride = out_stride;
			a.out_len = (uint32_t *)b->d_len.p;
			a.wm = workmem_bytes_power_of_two;
			a.lanes = g_compress_lanes;
			a.ctas_per_sm = g_ctas_per_sm;
			a.counter = (uint32_t *)b->d_ctr.p;
			TRY("compress launch", csb_launch_compress(&a, (csb_stream_t)b->s));
			TRY("pack launch", csb_launch_pack_stored((const uint8_t *)b->d_slots.p, out_stride, (const uint32_t *)b->d_len.p, nb,
								  (const uint8_t *)b->d_in.p, page_size, in_bytes, (uint32_t *)b->d_clen.p,
								  (uint8_t *)b->d_packed.p, (uint64_t *)b->d_off.p, (csb_stream_t)b->s));
			TRY("D2H total", cudaMemcpyAsync(b->h_res.p, (uint64_t *)b->d_off.p + nb, 8, cudaMemcpyDeviceToHost, b->s));
			TRY("D2H index", cudaMemcpyAsync(cont + 4 + 4 * done, b->d_clen.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, b->s));
			TRY("event record", cudaEventRecord(b->ev, b->s));
			done += nb;
			issued++;
		}
	}
	for (k = 0; k < BC_PIPE; k++)
		TRY("sync", cudaStreamSynchronize(BC[k].s));
	*container_length = payload_pos;
out:
	if (rc && bc_ready)
		for (k = 0; k < BC_PIPE; k++)
			cudaStreamSynchronize(BC[k].s);
	pthread_mutex_unlock(&C.mu);
	return rc;
}

int csnappy_bc_decompress_host(const void *h_container, uint64_t container_length, uint32_t page_size, void *h_out,
			       uint64_t out_capacity, uint64_t *out_length, uint32_t *failed_page)
{
	int rc = 0, k, first_err = 0;
	const uint8_t *cont = (const uint8_t *)h_container;
	uint32_t nr32 = 0, chunk, max_clen;
	uint64_t nr, done, ipos, issued = 0, retire_next = 0, produced_total = 0, err_page = 0;
	const uint32_t *idx;
	if (page_size == 0 || page_size > CSB_FRAGMENT_MAX)
		return set_err("page_size outside 1..32768", 0);
	if (!h_container || !out_length || container_length < 4)
		return set_err("container missing or shorter than its header", 0);
	memcpy(&nr32, cont, 4);
	nr = nr32;
	if (container_length < 4 + 4 * nr)
		return CSNAPPY_E_DATA_MALFORMED;
	if (out_capacity < nr * page_size || (nr && !h_out))
		return CSNAPPY_E_OUTPUT_INSUF;
	*out_length = 0;
	if (nr == 0)
		return 0;
	idx = (const uint32_t *)(cont + 4); /* (4-byte aligned if the container is) */
	ipos = 4 + 4 * nr;
	chunk = bc_chunk_pages(page_size, nr);
	max_clen = csnappy_max_compressed_length(page_size);

	pthread_mutex_lock(&C.mu);
	TRY("stream create", bc_init());
	for (k = 0; k < BC_PIPE; k++) {
		struct bc_slot *b = &BC[k];
		TRY("cudaMalloc(packed)", grow_dev(&b->d_packed, (size_t)chunk * max_clen + 64));
		TRY("cudaMalloc(clen)", grow_dev(&b->d_clen, (size_t)chunk * 4 + 64));
		TRY("cudaMalloc(off)", grow_dev(&b->d_off, ((size_t)chunk + 1) * 8 + 64));
		TRY("cudaMalloc(out)", grow_dev(&b->d_out, (size_t)chunk * page_size + 64));
		TRY("cudaMalloc(res)", grow_dev(&b->d_res, (size_t)chunk * 8 + 64));
		TRY("cudaMallocHost(res)", grow_pin(&b->h_res, (size_t)chunk * 8 + 64));
		TRY("cudaMalloc(ctr)", grow_dev(&b->d_ctr, 64));
		b->busy = 0;
	}
	for (done = 0; done < nr || retire_next < issued;) {
		if (retire_next < issued && (issued - retire_next >= BC_PIPE || done >= nr)) {
			struct bc_slot *b = &BC[retire_next % BC_PIPE];
			const uint32_t *olen = (const uint32_t *)b->h_res.p;
			const int32_t *st = (const int32_t *)b->h_res.p + b->n;
			uint32_t i;
			TRY("sync", cudaStreamSynchronize(b->s));
			for (i = 0; i < b->n; i++) {
				if (st[i] != 0 && (!first_err || b->first + i < err_page)) {
					first_err = st[i];
					err_page = b->first + i;
				}
				produced_total += olen[i];
			}
			retire_next++;
			continue;
		}
		{
			struct bc_slot *b = &BC[issued % BC_PIPE];
			uint64_t left_pages = nr - done, bytes = 0;
			uint32_t nb = left_pages < chunk ? (uint32_t)left_pages : chunk, i, longest = 0;
			struct csb_decompress_args a;
			for (i = 0; i < nb; i++) {
				uint32_t cl;
				memcpy(&cl, (const uint8_t *)idx + 4 * (done + i), 4);
				if (cl > max_clen || ipos + bytes + cl > container_length) {
					/* a size no writer produces, or a payload past the end of the container */
					if (!first_err) {
						first_err = CSNAPPY_E_DATA_MALFORMED;
						err_page = done + i;
					}
					nb = i;
					left_pages = 0;
					break;
				}
				bytes += cl;
				if (cl > longest)
					longest = cl;
			}
			if (nb) {
				TRY("H2D payload", cudaMemcpyAsync(b->d_packed.p, cont + ipos, bytes, cudaMemcpyHostToDevice, b->s));
				TRY("H2D index", cudaMemcpyAsync(b->d_clen.p, (const uint8_t *)idx + 4 * done, (size_t)nb * 4, cudaMemcpyHostToDevice, b->s));
				TRY("scan launch", csb_launch_scan((const uint32_t *)b->d_clen.p, nb, (uint64_t *)b->d_off.p, (csb_stream_t)b->s));
				memset(&a, 0, sizeof(a));
				a.in = (const uint8_t *)b->d_packed.p;
				a.in_off = (const uint64_t *)b->d_off.p;
				a.in_len = (const uint32_t *)b->d_clen.p;
				a.n_blocks = nb;
				a.out = (uint8_t *)b->d_out.p;
				a.out_stride = page_size;
				a.uniform_cap = page_size;
				a.out_len = (uint32_t *)b->d_res.p;
				a.status = (int32_t *)b->d_res.p + nb;
				a.flags = CSNAPPY_BATCH_RAW_IF_FULL;
				a.max_in_len = longest;
				a.counter = (uint32_t *)b->d_ctr.p;
				a.lanes = g_decompress_lanes;
				a.stage_input = g_stage_input;
				a.smem_kb = g_smem_kb;
				a.ctas_per_sm = g_ctas_per_sm;
				TRY("decompress launch", csb_launch_decompress(&a, (csb_stream_t)b->s));
				TRY("D2H pages", cudaMemcpyAsync((uint8_t *)h_out + done * page_size, b->d_out.p, (size_t)nb * page_size,
								 cudaMemcpyDeviceToHost, b->s));
				TRY("D2H result", cudaMemcpyAsync(b->h_res.p, b->d_res.p, (size_t)nb * 8, cudaMemcpyDeviceToHost, b->s));
				b->first = done;
				b->n = nb;
				ipos += bytes;
				done += nb;
				issued++;
			}
			if (left_pages == 0)
				done = nr; /* stop queueing after a malformed index entry */
		}
	}
	*out_length = produced_total;
	if (first_err) {
		if (failed_page)
			*failed_page = (uint32_t)err_page;
		rc = first_err;
	}
out:
	if (bc_ready)
		for (k = 0; k < BC_PIPE; k++)
			cudaStreamSynchronize(BC[k].s);
	pthread_mutex_unlock(&C.mu);
	return rc;
}

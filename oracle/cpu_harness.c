/*
 * cpu_harness.c -- pthread driver that runs a CPU Snappy block codec over a
 * strided batch of pages.  TEST / BASELINE INFRASTRUCTURE ONLY (see the header
 * of snappy_oracle.c): it is the `cpu_baseline` and `--impl reference` leg of
 * bench.py and the bulk checker of the GPU parity tests.
 *
 * impl = 0  our restatement (snappy_oracle.c)            -> "port"
 * impl = 1  the unmodified reference, dlopen()ed from
 *           oracle/_ref/libcsnappy_ref.so               -> "reference"
 *
 * Call shape follows the reference's own batch callers: one
 * csnappy_compress_fragment / csnappy_decompress_noheader per page with a
 * private working memory per thread (block_compressor.c:113-134,
 * kernel_3_2_10.patch:1348-1372).  Static contiguous partition of the page
 * range, CLOCK_MONOTONIC around the parallel region.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

uint32_t oracle_compress_fragment(const uint8_t *in, uint32_t n, uint8_t *out, int wm);
int oracle_decompress_noheader(const uint8_t *src, uint32_t src_len, uint8_t *dst, uint32_t *dst_len);

typedef char *(*ref_frag_fn)(const char *, const uint32_t, char *, void *, const int);
typedef int (*ref_dec_fn)(const char *, uint32_t, char *, uint32_t *);

static void *g_ref;
static ref_frag_fn g_ref_frag;
static ref_dec_fn g_ref_dec;

int harness_open_ref(const char *path)
{
	if (g_ref)
		return 0;
	g_ref = dlopen(path, RTLD_NOW | RTLD_LOCAL);
	if (!g_ref)
		return -1;
	g_ref_frag = (ref_frag_fn)dlsym(g_ref, "csnappy_compress_fragment");
	g_ref_dec = (ref_dec_fn)dlsym(g_ref, "csnappy_decompress_noheader");
	return (g_ref_frag && g_ref_dec) ? 0 : -2;
}

struct job {
	int impl, decode, wm;
	const uint8_t *in;
	uint64_t in_stride;
	const uint32_t *in_len; /* NULL => uniform_len */
	uint32_t uniform_len;
	uint8_t *out;
	uint64_t out_stride;
	uint32_t cap;
	uint32_t *out_len;
	int32_t *status;
	uint64_t first, last;
};

static void *worker(void *arg)
{
	struct job *j = (struct job *)arg;
	void *wm_buf = malloc((size_t)1 << 16);
	uint64_t i;
	for (i = j->first; i < j->last; i++) {
		const uint8_t *src = j->in + i * j->in_stride;
		uint8_t *dst = j->out + i * j->out_stride;
		uint32_t len = j->in_len ? j->in_len[i] : j->uniform_len;
		if (!j->decode) {
			if (j->impl)
				j->out_len[i] = (uint32_t)(g_ref_frag((const char *)src, len, (char *)dst,
								      wm_buf, j->wm) - (char *)dst);
			else
				j->out_len[i] = oracle_compress_fragment(src, len, dst, j->wm);
		} else {
			uint32_t olen = j->cap;
			int rc = j->impl ? g_ref_dec((const char *)src, len, (char *)dst, &olen)
					 : oracle_decompress_noheader(src, len, dst, &olen);
			j->status[i] = rc;
			j->out_len[i] = rc == 0 ? olen : 0;
		}
	}
	free(wm_buf);
	return NULL;
}

static double run(struct job *proto, uint64_t n, int nthreads)
{
	struct timespec t0, t1;
	pthread_t *tid;
	struct job *jobs;
	int t;
	if (nthreads < 1)
		nthreads = 1;
	if ((uint64_t)nthreads > n && n)
		nthreads = (int)n;
	tid = (pthread_t *)malloc(sizeof(*tid) * nthreads);
	jobs = (struct job *)malloc(sizeof(*jobs) * nthreads);
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (t = 0; t < nthreads; t++) {
		jobs[t] = *proto;
		jobs[t].first = n * t / nthreads;
		jobs[t].last = n * (t + 1) / nthreads;
		if (nthreads == 1)
			worker(&jobs[t]);
		else
			pthread_create(&tid[t], NULL, worker, &jobs[t]);
	}
	if (nthreads > 1)
		for (t = 0; t < nthreads; t++)
			pthread_join(tid[t], NULL);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	free(tid);
	free(jobs);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* returns seconds spent in the parallel region, < 0 on setup error */
double harness_compress_pages(int impl, const uint8_t *in, uint64_t in_stride,
			      const uint32_t *in_len, uint32_t uniform_len,
			      uint64_t n_pages, uint8_t *out, uint64_t out_stride,
			      uint32_t *out_len, int wm, int nthreads)
{
	struct job j;
	if (impl && !g_ref)
		return -1.0;
	memset(&j, 0, sizeof(j));
	j.impl = impl; j.decode = 0; j.wm = wm;
	j.in = in; j.in_stride = in_stride; j.in_len = in_len; j.uniform_len = uniform_len;
	j.out = out; j.out_stride = out_stride; j.out_len = out_len;
	return run(&j, n_pages, nthreads);
}

double harness_decompress_pages(int impl, const uint8_t *in, uint64_t in_stride,
				const uint32_t *in_len, uint64_t n_pages,
				uint8_t *out, uint64_t out_stride, uint32_t cap,
				uint32_t *out_len, int32_t *status, int nthreads)
{
	struct job j;
	if (impl && !g_ref)
		return -1.0;
	memset(&j, 0, sizeof(j));
	j.impl = impl; j.decode = 1;
	j.in = in; j.in_stride = in_stride; j.in_len = in_len;
	j.out = out; j.out_stride = out_stride; j.cap = cap;
	j.out_len = out_len; j.status = status;
	return run(&j, n_pages, nthreads);
}

/*
 * dropin_main.c -- a C caller built ONLY against include/ and linked against
 * libcsnappy_b200.so, the way cl_tester / block_compressor / zram use the reference
 * (cl_tester.c:14-114, 167-238; block_compressor.c:275-394; kernel_3_2_10.patch:1346-1376).
 *
 *   dropin_main <urls.10K> <urls.10K.snappy>
 *
 * Checks, through the plain C ABI: csnappy_compress(wm 15) == the fixture, round trip,
 * -2 / -3 / -5 codes of cl_tester's decompression self test, the zram macros on one page,
 * the batched zram helpers and the block_compressor container.  Prints "dropin ok".
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "csnappy.h"
#include "csnappy_batch.h"
#include "csnappy_zram.h"

#define CHECK(cond)                                                                     \
	do {                                                                            \
		if (!(cond)) {                                                          \
			fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
			return 1;                                                       \
		}                                                                       \
	} while (0)

static char *slurp(const char *path, uint32_t *len)
{
	FILE *f = fopen(path, "rb");
	char *buf;
	long n;
	if (!f)
		return NULL;
	fseek(f, 0, SEEK_END);
	n = ftell(f);
	fseek(f, 0, SEEK_SET);
	buf = malloc((size_t)n + 16);
	if (fread(buf, 1, (size_t)n, f) != (size_t)n)
		return NULL;
	fclose(f);
	*len = (uint32_t)n;
	return buf;
}

int main(int argc, char **argv)
{
	uint32_t n = 0, gn = 0, clen = 0, olen = 0, hdr = 0;
	char *in, *golden, *comp, *back;
	char workmem[1]; /* accepted and ignored by this library */
	static const char fake[] = "\x32\xc4\x66\x6f\x6f\x6f\x6f\x6f\x6f"; /* cl_tester.c:167 */
	CHECK(argc == 3);
	CHECK((in = slurp(argv[1], &n)) != NULL);
	CHECK((golden = slurp(argv[2], &gn)) != NULL);
	comp = malloc(csnappy_max_compressed_length(n));
	back = malloc(n + 16);

	/* cl_tester -c / -d, at the table size that reproduces the checked-in fixture */
	csnappy_compress(in, n, comp, &clen, workmem, 15);
	CHECK(clen == gn && memcmp(comp, golden, gn) == 0);
	CHECK(csnappy_get_uncompressed_length(comp, clen, &olen) == 3 && olen == n);
	CHECK(csnappy_decompress(comp, clen, back, n) == CSNAPPY_E_OK && memcmp(back, in, n) == 0);

	/* cl_tester -S d (cl_tester.c:167-238) */
	CHECK(csnappy_decompress(comp, clen, back, n - 1) == CSNAPPY_E_OUTPUT_INSUF);
	hdr = 3;
	olen = 4096;
	CHECK(csnappy_decompress_noheader(comp + hdr, clen - hdr, back, &olen) == CSNAPPY_E_OUTPUT_OVERRUN && olen == 4096);
	CHECK(csnappy_decompress(fake, 9, back, 50) == CSNAPPY_E_DATA_MALFORMED);
	olen = 50;
	CHECK(csnappy_decompress_noheader(fake + 1, 8, back, &olen) == CSNAPPY_E_DATA_MALFORMED);

	/* zram: one page through the macros, then a batch of pages */
	{
		unsigned char slot[4816], page[4096];
		size_t zl = 0, pl = sizeof(page);
		CHECK(COMPRESS((const unsigned char *)in, 4096, slot, &zl, workmem) == 0 && zl > 0 && zl < 4096);
		CHECK(DECOMPRESS(slot, zl, page, &pl) == 0 && pl == 4096 && memcmp(page, in, 4096) == 0);
	}
	{
		uint32_t pages = n / 4096, i, *zlen = malloc(4 * pages), *plen = malloc(4 * pages);
		int32_t *st = malloc(4 * pages);
		unsigned char *slots = malloc((size_t)pages * 4816);
		CHECK(csnappy_zram_compress_pages((const unsigned char *)in, pages, slots, 4816, zlen) == 0);
		memset(back, 0, n);
		CHECK(csnappy_zram_decompress_pages(slots, 4816, zlen, pages, (unsigned char *)back, plen, st) == 0);
		for (i = 0; i < pages; i++)
			CHECK(st[i] == 0 && plen[i] == 4096);
		CHECK(memcmp(back, in, (size_t)pages * 4096) == 0);
	}

	/* block_compressor container (block_compressor.c:275-394) */
	{
		uint64_t cap = csnappy_bc_max_container_length(n, 4096), cl = 0, ol = 0;
		char *cont = malloc(cap);
		uint32_t nr_pages = 0;
		CHECK(csnappy_bc_compress_host(in, n, 4096, cont, cap, &cl, 13) == 0);
		memcpy(&nr_pages, cont, 4);
		CHECK(nr_pages == (n + 4095) / 4096 && cl < n);
		memset(back, 0, n);
		back = realloc(back, (size_t)nr_pages * 4096);
		CHECK(csnappy_bc_decompress_host(cont, cl, 4096, back, (uint64_t)nr_pages * 4096, &ol, NULL) == 0);
		CHECK(ol == n && memcmp(back, in, n) == 0);
	}
	/* the same container over every visible device from this one process, and over device 0 listed
	 * twice (two workers, two contexts: the chunk-position hand-over without needing two GPUs) */
	{
		uint64_t cap = csnappy_bc_max_container_length(n, 4096), cl = 0, cl2 = 0, cl3 = 0, ol = 0;
		char *c1 = malloc(cap), *c2 = malloc(cap), *c3 = malloc(cap);
		const int twice[2] = {0, 0};
		uint32_t nr_pages = (n + 4095) / 4096;
		CHECK(csnappy_b200_device_count() >= 1);
		CHECK(csnappy_bc_compress_host(in, n, 4096, c1, cap, &cl, 13) == 0);
		CHECK(csnappy_bc_compress_host_multi(in, n, 4096, c2, cap, &cl2, 13, NULL, 0) == 0);
		CHECK(csnappy_bc_compress_host_multi(in, n, 4096, c3, cap, &cl3, 13, twice, 2) == 0);
		CHECK(cl2 == cl && memcmp(c1, c2, cl) == 0);
		CHECK(cl3 == cl && memcmp(c1, c3, cl) == 0);
		memset(back, 0, (size_t)nr_pages * 4096);
		CHECK(csnappy_bc_decompress_host_multi(c2, cl2, 4096, back, (uint64_t)nr_pages * 4096, &ol, NULL, NULL, 0) == 0);
		CHECK(ol == n && memcmp(back, in, n) == 0);
		memset(back, 0, (size_t)nr_pages * 4096);
		CHECK(csnappy_bc_decompress_host_multi(c3, cl3, 4096, back, (uint64_t)nr_pages * 4096, &ol, NULL, twice, 2) == 0);
		CHECK(ol == n && memcmp(back, in, n) == 0);
	}
	printf("dropin ok: %u -> %u bytes, %llu kernels launched\n", n, clen, (unsigned long long)csnappy_b200_kernel_launches());
	return 0;
}

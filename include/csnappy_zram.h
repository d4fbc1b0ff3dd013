/*
 * csnappy_zram.h -- the zram glue of the reference, over libcsnappy_b200.so.
 *
 * The reference's Linux 3.2 zram patch wraps the hot path in two helpers and
 * two macros (/root/reference/kernel_3_2_10.patch:1346-1376):
 *     COMPRESS(s, sl, d, dl, wm)    -> snappy_compress_(...)   one PAGE_SIZE fragment,
 *                                      table of 1 << (PAGE_SHIFT + 1) bytes, no header
 *     DECOMPRESS(s, sl, d, dl)      -> snappy_decompress_(...) raw stream into one page
 * This header keeps those names, argument orders and size_t lengths (per page, host
 * pointers) and adds the shape that suits a GPU: the same two operations over a BATCH
 * of pages, which is what a write-back / swap-out path hands over at once.
 * The policy "store the page uncompressed if clen > max_zpage_size" stays with the
 * caller, exactly as in the patch (:1425-1431).
 */
#ifndef CSNAPPY_B200_CSNAPPY_ZRAM_H_
#define CSNAPPY_B200_CSNAPPY_ZRAM_H_

#include <stddef.h>
#include <stdint.h>

#include "csnappy.h"
#include "csnappy_batch.h"

#ifndef CSNAPPY_ZRAM_PAGE_SHIFT
#define CSNAPPY_ZRAM_PAGE_SHIFT 12
#endif
#define CSNAPPY_ZRAM_PAGE_SIZE (1u << CSNAPPY_ZRAM_PAGE_SHIFT)
/* kernel_3_2_10.patch:1346-1347 */
#define CSNAPPY_ZRAM_WMSIZE_ORDER ((CSNAPPY_ZRAM_PAGE_SHIFT > 14) ? (15) : (CSNAPPY_ZRAM_PAGE_SHIFT + 1))
#define CSNAPPY_ZRAM_WMSIZE (1 << CSNAPPY_ZRAM_WMSIZE_ORDER)

/* kernel_3_2_10.patch:1348-1360 */
static inline int csnappy_zram_compress(const unsigned char *src, size_t src_len, unsigned char *dst,
					size_t *dst_len, void *workmem)
{
	const char *end = csnappy_compress_fragment((const char *)src, (uint32_t)src_len, (char *)dst, workmem,
						    CSNAPPY_ZRAM_WMSIZE_ORDER);
	*dst_len = (size_t)(end - (const char *)dst);
	return 0;
}

/* kernel_3_2_10.patch:1361-1372 */
static inline int csnappy_zram_decompress(const unsigned char *src, size_t src_len, unsigned char *dst,
					  size_t *dst_len)
{
	uint32_t dst_len_ = (uint32_t)*dst_len;
	int ret = csnappy_decompress_noheader((const char *)src, (uint32_t)src_len, (char *)dst, &dst_len_);
	*dst_len = (size_t)dst_len_;
	return ret;
}

#ifndef COMPRESS
#define COMPRESS(s, sl, d, dl, wm) csnappy_zram_compress(s, sl, d, dl, wm)
#define DECOMPRESS(s, sl, d, dl) csnappy_zram_decompress(s, sl, d, dl)
#endif

/*
 * Batched form: n_pages pages of CSNAPPY_ZRAM_PAGE_SIZE bytes at `pages` (contiguous), each
 * compressed into its own slot of slot_stride >= csnappy_max_compressed_length(PAGE_SIZE) bytes;
 * clen[i] receives the compressed size.  One call, pipelined through the device.
 */
static inline int csnappy_zram_compress_pages(const unsigned char *pages, uint32_t n_pages, unsigned char *slots,
					      uint64_t slot_stride, uint32_t *clen)
{
	return csnappy_batch_compress_fragments_host(pages, CSNAPPY_ZRAM_PAGE_SIZE, CSNAPPY_ZRAM_PAGE_SIZE, n_pages,
						     slots, slot_stride, clen, CSNAPPY_ZRAM_WMSIZE_ORDER);
}

/* Inverse: status[i] is the per-page return code of DECOMPRESS, out_len[i] the bytes produced. */
static inline int csnappy_zram_decompress_pages(const unsigned char *slots, uint64_t slot_stride,
						const uint32_t *clen, uint32_t n_pages, unsigned char *pages,
						uint32_t *out_len, int32_t *status)
{
	return csnappy_batch_decompress_host(slots, slot_stride, clen, n_pages, pages, CSNAPPY_ZRAM_PAGE_SIZE,
					     CSNAPPY_ZRAM_PAGE_SIZE, out_len, status, 0);
}

#endif /* CSNAPPY_B200_CSNAPPY_ZRAM_H_ */

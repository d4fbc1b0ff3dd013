/*
 * csnappy.h -- drop-in C API of the B200-native Snappy block codec.
 *
 * Same six symbols, argument order, types, macro values and error codes as the
 * reference header (/root/reference/csnappy.h:11-14, 30-129), so a caller built
 * against the reference relinks against libcsnappy_b200.so unchanged.  The
 * codec work behind every entry point runs in hand-written sm_100a CUDA
 * kernels; there is no CPU codec in this library.  All pointers in THIS header
 * are HOST pointers (the library stages them through device memory); the
 * device-resident batched entry points live in csnappy_batch.h.
 *
 * Differences a caller can observe:
 *   - working_memory is accepted and ignored: the hash table lives in shared
 *     memory, sized by workmem_bytes_power_of_two exactly as the reference
 *     sizes its table, so the emitted bytes are identical.
 *   - 9 <= workmem_bytes_power_of_two <= 16 is enforced (the reference
 *     documents 9..15 and ships 16 as default, csnappy.h:13,41).
 *   - a machine without a usable CUDA device makes the decompress calls return
 *     CSNAPPY_E_DEVICE and the compress calls (which have no error channel in
 *     this ABI) print a diagnostic and abort().
 *   - a copy tag or literal-length field cut off by end of input is defined as
 *     CSNAPPY_E_DATA_MALFORMED (undefined behaviour in the reference's x86
 *     path, csnappy_decompress.c:331-350).
 */
#ifndef CSNAPPY_B200_CSNAPPY_H_
#define CSNAPPY_B200_CSNAPPY_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference csnappy.h:11-14 */
#define CSNAPPY_VERSION 5
#define CSNAPPY_WORKMEM_BYTES_POWER_OF_TWO 16
#define CSNAPPY_WORKMEM_BYTES (1 << CSNAPPY_WORKMEM_BYTES_POWER_OF_TWO)

/* reference csnappy.h:124-129 */
#define CSNAPPY_E_OK 0
#define CSNAPPY_E_HEADER_BAD (-1)
#define CSNAPPY_E_OUTPUT_INSUF (-2)
#define CSNAPPY_E_OUTPUT_OVERRUN (-3)
#define CSNAPPY_E_INPUT_NOT_CONSUMED (-4) /* defined, never returned (as in the reference) */
#define CSNAPPY_E_DATA_MALFORMED (-5)
/* additions of this library (never produced by the reference) */
#define CSNAPPY_E_DEVICE (-100)  /* CUDA device / runtime / launch failure */
#define CSNAPPY_E_BAD_ARG (-101) /* argument outside the documented domain */

/*
 * Upper bound of the compressed size of source_len input bytes: 32 + n + n/6.
 * Replaces csnappy.h:30-31 (csnappy_compress.c:612-616).  Pure host arithmetic.
 */
uint32_t csnappy_max_compressed_length(uint32_t source_len)
#ifdef __GNUC__
	__attribute__((const))
#endif
	;

/*
 * Compress one fragment (input_length <= 32768) WITHOUT the length prefix.
 * output must hold csnappy_max_compressed_length(input_length) bytes.
 * Returns the end pointer into output.  Bytes written are identical to the
 * reference's for the same input and workmem_bytes_power_of_two.
 * Replaces csnappy.h:46-52 (csnappy_compress.c:469-606).
 */
char *csnappy_compress_fragment(const char *input, const uint32_t input_length,
				char *output, void *working_memory,
				const int workmem_bytes_power_of_two);

/*
 * Compress a whole buffer: varint32 length prefix, then one fragment per
 * 32 KiB chunk, the last (short) chunk using the reference's smaller-table
 * rule.  Fragments are compressed in parallel on the device and packed.
 * Replaces csnappy.h:65-72 (csnappy_compress.c:621-656).
 */
void csnappy_compress(const char *input, uint32_t input_length, char *compressed,
		      uint32_t *out_compressed_length, void *working_memory,
		      const int workmem_bytes_power_of_two);

/*
 * Parse the varint32 length prefix.  Returns bytes consumed (1..5) or
 * CSNAPPY_E_HEADER_BAD; *result is written even on error (partial value).
 * Replaces csnappy.h:83-87 (csnappy_decompress.c:45-71).  Pure host arithmetic.
 */
int csnappy_get_uncompressed_length(const char *start, uint32_t n, uint32_t *result);

/*
 * Decompress a stream WITH length prefix into dst[0..dst_len).
 * -1 bad header, -2 header length > dst_len, else the result of the raw
 * decoder with capacity = header length (a stream that ends early is OK).
 * Replaces csnappy.h:99-104 (csnappy_decompress.c:394-411).
 */
int csnappy_decompress(const char *src, uint32_t src_len, char *dst, uint32_t dst_len);

/*
 * Decompress a raw tag stream.  In: *dst_len = capacity.  Out (success only):
 * *dst_len = bytes produced.  0 / CSNAPPY_E_OUTPUT_OVERRUN / CSNAPPY_E_DATA_MALFORMED.
 * Replaces csnappy.h:114-119 (csnappy_decompress.c:319-387).
 */
int csnappy_decompress_noheader(const char *src, uint32_t src_len, char *dst,
				uint32_t *dst_len);

#ifdef __cplusplus
}
#endif
#endif /* CSNAPPY_B200_CSNAPPY_H_ */

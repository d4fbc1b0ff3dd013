// stream_kernel.cu -- placeholder until the parallel single-stream decoder lands (next commit)
#include "device_common.cuh"
#include "kernels.h"
extern "C" size_t csb_stream_aux_bytes(uint32_t, uint32_t, int) { return 64; }
extern "C" int csb_launch_decompress_stream(const uint8_t *, uint32_t, uint8_t *, uint32_t, uint32_t *, int32_t *, void *, void *,
					    csb_stream_t)
{
	return 1;
}

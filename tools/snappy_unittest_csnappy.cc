// snappy_unittest_csnappy.cc -- the CSNAPPY case of upstream's snappy_unittest benchmark harness, as the
// reference adds it (/root/reference/snappy_tester.patch:72-107, 128-135), over libcsnappy_b200.so:
//   Compress()    csnappy_compress(input, size, out, &destlen, workmem, CSNAPPY_WORKMEM_BYTES_POWER_OF_TWO)
//                 + CHECK_LE(destlen, csnappy_max_compressed_length(size))            (:72-89)
//   Uncompress()  csnappy_decompress(compressed, csize, out, size) == CSNAPPY_E_OK    (:96-107)
//   Measure()     blocks of 1 MiB ("[b 1M]", :133), every block compressed / uncompressed `repeats` times,
//                 the MEDIAN run reported, same output line as upstream's harness (:120-127):
//   CSNAPPY  [b 1M] bytes 702087 -> 357267 50.9%  comp 240.1 MB/s  uncomp 645.5 MB/s
// Build (tools/run_unittest_adapter.sh):  g++ -O2 -Iinclude tools/snappy_unittest_csnappy.cc -Lcsnappy_b200 -lcsnappy_b200
// Usage: snappy_unittest_csnappy [--repeats N] [--wm P] [--expect file.snappy] file...
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "csnappy.h"

#define CHECK(c)                                                                     \
	do {                                                                         \
		if (!(c)) {                                                          \
			fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
			exit(1);                                                     \
		}                                                                    \
	} while (0)

static std::string slurp(const char *path)
{
	FILE *f = fopen(path, "rb");
	CHECK(f != nullptr);
	std::string s;
	char buf[1 << 16];
	size_t n;
	while ((n = fread(buf, 1, sizeof(buf), f)) > 0)
		s.append(buf, n);
	fclose(f);
	return s;
}

static double now()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
	int repeats = 20, wm = CSNAPPY_WORKMEM_BYTES_POWER_OF_TWO;
	const size_t block_size = 1024 << 10;
	const char *expect = nullptr;
	std::vector<const char *> files;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--repeats") && i + 1 < argc)
			repeats = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--wm") && i + 1 < argc)
			wm = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--expect") && i + 1 < argc)
			expect = argv[++i];
		else
			files.push_back(argv[i]);
	}
	CHECK(!files.empty());
	for (const char *fname : files) {
		const std::string input = slurp(fname);
		const size_t nblocks = (input.size() + block_size - 1) / block_size;
		std::vector<std::string> comp(nblocks), back(nblocks);
		std::vector<double> ctime(repeats), utime(repeats);
		char *mem = new char[CSNAPPY_WORKMEM_BYTES];
		size_t csize = 0;
		for (int r = -1; r < repeats; r++) {  // r == -1: warm-up (first call creates the device context)
			double t0 = now();
			csize = 0;
			for (size_t b = 0; b < nblocks; b++) {
				const size_t at = b * block_size, n = std::min(block_size, input.size() - at);
				uint32_t destlen = 0;
				comp[b].resize(csnappy_max_compressed_length((uint32_t)n));
				csnappy_compress(input.data() + at, (uint32_t)n, &comp[b][0], &destlen, mem, wm);
				CHECK(destlen <= csnappy_max_compressed_length((uint32_t)n));
				comp[b].resize(destlen);
				csize += destlen;
			}
			double t1 = now();
			for (size_t b = 0; b < nblocks; b++) {
				const size_t at = b * block_size, n = std::min(block_size, input.size() - at);
				back[b].resize(n);
				CHECK(csnappy_decompress(comp[b].data(), (uint32_t)comp[b].size(), &back[b][0], (uint32_t)n) == CSNAPPY_E_OK);
			}
			double t2 = now();
			if (r >= 0) {
				ctime[r] = t1 - t0;
				utime[r] = t2 - t1;
			}
		}
		delete[] mem;
		for (size_t b = 0; b < nblocks; b++)
			CHECK(memcmp(back[b].data(), input.data() + b * block_size, back[b].size()) == 0);
		if (expect && nblocks == 1) {
			const std::string want = slurp(expect);
			CHECK(want == comp[0]);
			fprintf(stderr, "compressed bytes identical to %s\n", expect);
		}
		std::sort(ctime.begin(), ctime.end());
		std::sort(utime.begin(), utime.end());
		const double cm = ctime[repeats / 2], um = utime[repeats / 2];
		printf("%-8s [b %dM] bytes %6d -> %6d %4.1f%%  comp %5.1f MB/s  uncomp %5.1f MB/s\n", "CSNAPPY", (int)(block_size >> 20),
		       (int)input.size(), (int)csize, 100.0 * csize / std::max<size_t>(1, input.size()),
		       input.size() / cm / 1048576.0, input.size() / um / 1048576.0);
	}
	return 0;
}

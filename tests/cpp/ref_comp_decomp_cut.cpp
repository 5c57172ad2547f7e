// The reference's own round-trip / bounds test (tests/tests_comp_decomp.cpp) against the boundary, cut to the domain of
// the device path: element sizes 2, 4, 8 and levels 0..1 (open decision 2).
//
// The reference's translation unit is #included from /root/reference at build time -- nothing is copied: its
// test_vector() IS the check (sentinels behind dst and behind the decoded buffer, "an error implies dst_size <
// stenos_bound", round trip), its generators make the data.  Only the outer sweep is ours: the reference's sweeps
// 15 element sizes x 6 levels x 8 threads and cannot be restricted from outside.
// Built by oracle/Makefile (target reftests) once against the reference and once against libstenos_b200.so.
#define tests_comp_decomp reference_full_sweep_not_called
#include "tests/tests_comp_decomp.cpp"
#undef tests_comp_decomp

#include <cstdio>

static size_t g_calls = 0;

template<size_t N>
static void sweep(const char* dist_name)
{
	using type = std::array<char, N>;
	// sizes in elements: empty, tiny (Zstd tail), around one block, around one superblock, several superblocks.  No exact
	// multiple of the superblock: the reference's own decoder rejects those and this sweep must
	// pass against both libraries.
	const size_t sizes[] = { 0, 1, 37, 255, 256, 257, 4097, 131072 / N - 1, 131072 / N + 1, 131072 / N + 300, 3 * 131072 / N + 77, 700001 };
	for (size_t n : sizes) {
		std::vector<type> vec;
		if (strcmp("random", dist_name) == 0)
			vec = generate_random<type>(n);
		else if (strcmp("sorted", dist_name) == 0)
			vec = generate_random_sorted<type>(n);
		else
			vec = generate_same<type>(n);
		const size_t bytes = vec.size() * sizeof(type);
		std::mt19937 rng((unsigned)(n + N));
		std::uniform_int_distribution<int> step(0, (int)(bytes > 100 ? bytes / 10 : 10));
		for (int level = 0; level <= 1; ++level) {
			long long dst_size = (long long)stenos_bound(bytes);
			for (;;) { // the reference's shrinking dst_size loop (tests_comp_decomp.cpp:158-170)
				test_vector(vec, dist_name, level, 1, (size_t)dst_size);
				++g_calls;
				if (dst_size == 0)
					break;
				dst_size -= step(rng);
				if (dst_size < 0)
					dst_size = 0;
			}
		}
	}
}

int main()
{
	for (const char* d : { "same", "sorted", "random" }) {
		sweep<2>(d);
		sweep<4>(d);
		sweep<8>(d);
	}
	printf("ref_comp_decomp_cut ok: %zu test_vector calls\n", g_calls);
	return 0;
}

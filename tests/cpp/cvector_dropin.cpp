// Drop-in check of the boundary: the reference's UNCHANGED header-only container
// stenos/cvector.hpp, compiled once against the reference library and once against libstenos_b200.so, must behave
// the same and serialize to the same bytes.  Usage: cvector_dropin <out-file>; prints a summary line.
// Built by oracle/Makefile (target cvector) into oracle/_ref/ -- the header comes from /root/reference, nothing is copied.
#include "stenos/cvector.hpp"

#include <cstdint>
#include <cstdio>
#include <vector>

template<class T>
static bool run(size_t n, std::vector<char>& out, uint64_t& checksum)
{
	stenos::cvector<T> v;
	std::vector<T> ref;
	uint64_t x = 88172645463325252ull;
	for (size_t i = 0; i < n; ++i) {
		x ^= x << 13;
		x ^= x >> 7;
		x ^= x << 17;
		// ramp + noise, runs, and a stretch of incompressible values
		T val = (T)(3 * i + (x & 15));
		if ((i % 5000) > 4000)
			val = (T)42;
		if (i > n / 2 && i < n / 2 + 3000)
			val = (T)x;
		v.push_back(val);
		ref.push_back(val);
	}
	// random access writes (dirty buckets are re-compressed), then a full read back
	for (size_t i = 0; i < n; i += 997) {
		v[i] = (T)(ref[i] + 1);
		ref[i] = (T)(ref[i] + 1);
	}
	v.shrink_to_fit();
	bool ok = v.size() == ref.size();
	size_t i = 0;
	for (auto it = v.begin(); it != v.end() && ok; ++it, ++i) {
		const T got = *it;
		ok = got == ref[i];
		checksum = checksum * 1099511628211ull + (uint64_t)got;
	}
	std::vector<char> buf(n * sizeof(T) + n / 16 + 4096);
	const size_t r = v.serialize(buf.data(), buf.size());
	if (stenos_has_error(r))
		return false;
	out.insert(out.end(), buf.begin(), buf.begin() + r);
	// and back
	stenos::cvector<T> w;
	const size_t d = w.deserialize(buf.data(), r);
	ok = ok && !stenos_has_error(d) && w.size() == ref.size();
	for (size_t k = 0; k < ref.size() && ok; k += 13)
		ok = (T)w[k] == ref[k];
	return ok;
}

int main(int argc, char** argv)
{
	std::vector<char> out;
	uint64_t checksum = 1469598103934665603ull;
	const bool ok = run<int32_t>(100000, out, checksum) && run<int64_t>(30000, out, checksum) && run<int16_t>(50000, out, checksum);
	if (argc > 1) {
		FILE* f = std::fopen(argv[1], "wb");
		if (!f)
			return 2;
		std::fwrite(out.data(), 1, out.size(), f);
		std::fclose(f);
	}
	std::printf("ok=%d serialized=%zu checksum=%016llx\n", ok ? 1 : 0, out.size(), (unsigned long long)checksum);
	return ok ? 0 : 1;
}

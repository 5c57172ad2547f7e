// The reference's own container test (tests/test_cvector.cpp: cvector<size_t>, cvector<int>, move-only and atomic
// payloads -- all element sizes 4 and 8) against the boundary.  The translation unit is #included from /root/reference at
// build time, unchanged, nothing is copied; this file only supplies main().  Built by oracle/Makefile (target reftests).
#include "tests/test_cvector.cpp"

int main(int argc, char** argv)
{
	const int r = test_cvector(argc, argv);
	printf("ref_test_cvector %s\n", r == 0 ? "ok" : "FAILED");
	return r;
}

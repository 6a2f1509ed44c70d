// Dumps the reference's OpenCL C program text (src/kernel.hpp:7-18 get_opencl_c_code(), defined in the
// translation unit of src/kernel.cpp) to a file. Built and run only where the reference tree is mounted;
// the output goes to oracle/_ref/ (git-ignored) and is never committed.
#include <cstdio>
#include <string>
std::string get_opencl_c_code();
int main(int argc, char** argv) {
	if(argc<2) { std::fprintf(stderr, "usage: %s out.cl\n", argv[0]); return 2; }
	const std::string code = get_opencl_c_code();
	std::FILE* f = std::fopen(argv[1], "wb");
	if(!f) return 1;
	std::fwrite(code.data(), 1, code.size(), f);
	std::fclose(f);
	return 0;
}

// Prints the reference's own to_string(float) (src/utilities.hpp:2745-2754) for each float given as a hex bit
// pattern on stdin, one per line. Used only by tests (in this container) to pin orc_float_to_string().
#include "utilities.hpp"
#include <cstring>
#include <cstdio>
int main() {
	unsigned int bits;
	while(std::scanf("%x", &bits)==1) {
		float x; std::memcpy(&x, &bits, 4);
		std::printf("%s\n", to_string(x).c_str());
	}
	return 0;
}

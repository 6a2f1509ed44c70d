// main.cpp -- console driver: the scene runs on a compute thread while the main thread refreshes the metrics line,
// as in the reference's non-graphics main() (FluidX3D v3.7 src/main.cpp:146-164). Arguments are CUDA device IDs, one per domain.
#include "info.hpp"
#include "lbm.hpp"
#include "setup.hpp"
#include <atomic>

vector<string> main_arguments;
std::atomic<bool> running(true);

int main(int argc, char* argv[]) {
	main_arguments = vector<string>(argv+1, argv+argc);
	info.print_logo();
	thread compute_thread([]() { main_setup(); running = false; });
	while(running) { info.print_update(); sleep(0.050); }
	compute_thread.join();
	return 0;
}

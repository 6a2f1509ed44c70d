// emul_runtime.cpp -- TEST INFRASTRUCTURE ONLY. Host stand-ins for the device/stream/memory/IPC/rendezvous half of the
// C ABI (the part fx3d_runtime.cu implements with CUDA), so that the product's host logic (fluidx3d_b200/lbm.py) and the
// kernel sources compiled against cuda_emul.hpp can be exercised end to end on a machine without a GPU, including the
// one-process-per-domain path (buffers are POSIX shared memory, so "IPC handles" work across processes and the
// rendezvous counters really synchronise two processes). Linked only into tests/_build/libfx3d_emul.so.
#include "../../include/fx3d.h"
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <sched.h>

namespace fx3d { void set_error(const std::string& msg); }

namespace {
struct Block { std::string name; size_t bytes; };
std::mutex g_mu;
std::map<void*, Block> g_blocks;     // my allocations
std::map<void*, size_t> g_opened;    // mappings of other processes' allocations
std::atomic<uint64_t> g_counter{0};
}

extern "C" {

int fx3d_device_count(int* count) { if(!count) return FX3D_ERR_INVALID; *count = 1; return FX3D_OK; }
int fx3d_device_get_info(int device, fx3d_device_info* info) {
	if(!info) return FX3D_ERR_INVALID;
	std::memset(info, 0, sizeof(*info));
	std::snprintf(info->name, sizeof(info->name), "host emulation (tests only)");
	info->id = device; info->sm_count = 1; info->clock_mhz = 1000; info->memory_bytes = 1ull<<34;
	return FX3D_OK;
}
int fx3d_device_enable_peer(int, int) { return FX3D_OK; }
int fx3d_device_sync(int) { return FX3D_OK; }
int fx3d_stream_create(int, fx3d_stream* stream) { if(!stream) return FX3D_ERR_INVALID; *stream = reinterpret_cast<fx3d_stream>(new int(0)); return FX3D_OK; }
int fx3d_stream_destroy(int, fx3d_stream stream) { delete reinterpret_cast<int*>(stream); return FX3D_OK; }
int fx3d_stream_sync(int, fx3d_stream) { return FX3D_OK; }
int fx3d_event_create(int, fx3d_event* event) { if(!event) return FX3D_ERR_INVALID; *event = reinterpret_cast<fx3d_event>(new double(0.0)); return FX3D_OK; }
int fx3d_event_destroy(int, fx3d_event event) { delete reinterpret_cast<double*>(event); return FX3D_OK; }
int fx3d_event_record(int, fx3d_event event, fx3d_stream) {
	*reinterpret_cast<double*>(event) = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
	return FX3D_OK;
}
int fx3d_event_sync(int, fx3d_event) { return FX3D_OK; }
int fx3d_event_elapsed_ms(fx3d_event a, fx3d_event b, float* ms) { if(!ms) return FX3D_ERR_INVALID; *ms = (float)(*reinterpret_cast<double*>(b)-*reinterpret_cast<double*>(a)); return FX3D_OK; }
int fx3d_stream_wait_event(int, fx3d_stream, fx3d_event) { return FX3D_OK; }

int fx3d_malloc(int, size_t bytes, void** ptr) {
	if(!ptr) return FX3D_ERR_INVALID;
	const size_t n = ((bytes ? bytes : 1u)+256u+4095u)&~(size_t)4095u; // 256 bytes of slack like the product's fx3d_malloc
	char name[64];
	std::snprintf(name, sizeof(name), "/fx3d_emul_%d_%llu", (int)getpid(), (unsigned long long)g_counter++);
	const int fd = shm_open(name, O_CREAT|O_EXCL|O_RDWR, 0600);
	if(fd<0 || ftruncate(fd, (off_t)n)!=0) { if(fd>=0) { close(fd); shm_unlink(name); } fx3d::set_error("emulation: shm_open failed"); return FX3D_ERR_OUT_OF_MEMORY; }
	void* p = mmap(nullptr, n, PROT_READ|PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if(p==MAP_FAILED) { shm_unlink(name); fx3d::set_error("emulation: mmap failed"); return FX3D_ERR_OUT_OF_MEMORY; }
	std::memset(p, 0, n);
	std::lock_guard<std::mutex> lock(g_mu);
	g_blocks[p] = Block{ name, n };
	*ptr = p;
	return FX3D_OK;
}
int fx3d_free(int, void* ptr) {
	if(!ptr) return FX3D_OK;
	std::lock_guard<std::mutex> lock(g_mu);
	auto it = g_blocks.find(ptr);
	if(it==g_blocks.end()) { fx3d::set_error("emulation: free of unknown pointer"); return FX3D_ERR_INVALID; }
	munmap(ptr, it->second.bytes);
	shm_unlink(it->second.name.c_str());
	g_blocks.erase(it);
	return FX3D_OK;
}
int fx3d_host_alloc(size_t bytes, void** ptr) { if(!ptr) return FX3D_ERR_INVALID; *ptr = std::aligned_alloc(4096, ((bytes ? bytes : 1u)+4095u)&~(size_t)4095u); return *ptr ? FX3D_OK : FX3D_ERR_OUT_OF_MEMORY; }
int fx3d_host_free(void* ptr) { std::free(ptr); return FX3D_OK; }
int fx3d_memcpy_h2d(int, void* dst, const void* src, size_t bytes, fx3d_stream, int) { if(bytes) std::memcpy(dst, src, bytes); return FX3D_OK; }
int fx3d_memcpy_d2h(int, void* dst, const void* src, size_t bytes, fx3d_stream, int) { if(bytes) std::memcpy(dst, src, bytes); return FX3D_OK; }
int fx3d_memset(int, void* dst, int value, size_t bytes, fx3d_stream) { if(bytes) std::memset(dst, value, bytes); return FX3D_OK; }
int fx3d_fill_f32(int, float* dst, float value, size_t count, fx3d_stream) { for(size_t k=0u; k<count; k++) dst[k] = value; return FX3D_OK; }

int fx3d_ipc_get_handle(int, void* ptr, void* handle64) {
	std::lock_guard<std::mutex> lock(g_mu);
	auto it = g_blocks.find(ptr);
	if(it==g_blocks.end()) { fx3d::set_error("emulation: ipc handle of unknown pointer"); return FX3D_ERR_INVALID; }
	std::memset(handle64, 0, 64);
	std::memcpy(handle64, &it->second.bytes, sizeof(size_t));
	std::strncpy((char*)handle64+8, it->second.name.c_str(), 55);
	return FX3D_OK;
}
int fx3d_ipc_open_handle(int, const void* handle64, void** ptr) {
	size_t n; std::memcpy(&n, handle64, sizeof(size_t));
	const int fd = shm_open((const char*)handle64+8, O_RDWR, 0600);
	if(fd<0) { fx3d::set_error("emulation: cannot open peer buffer"); return FX3D_ERR_INVALID; }
	void* p = mmap(nullptr, n, PROT_READ|PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if(p==MAP_FAILED) return FX3D_ERR_OUT_OF_MEMORY;
	std::lock_guard<std::mutex> lock(g_mu);
	g_opened[p] = n; *ptr = p;
	return FX3D_OK;
}
int fx3d_ipc_close_handle(int, void* ptr) {
	std::lock_guard<std::mutex> lock(g_mu);
	auto it = g_opened.find(ptr);
	if(it==g_opened.end()) return FX3D_ERR_INVALID;
	munmap(ptr, it->second); g_opened.erase(it);
	return FX3D_OK;
}

int fx3d_rendezvous_signal(int, uint64_t* const* peer_arrays, int n_peers, int my_index, uint64_t value, fx3d_stream) {
	for(int k=0; k<n_peers; k++) __atomic_store_n(peer_arrays[k]+my_index, value, __ATOMIC_RELEASE);
	return FX3D_OK;
}
int fx3d_rendezvous_wait(int, uint64_t* my_array, const int* peer_indices, int n_peers, uint64_t value, int timeout_ms, fx3d_stream) {
	const auto start = std::chrono::steady_clock::now();
	for(int k=0; k<n_peers; k++) {
		while(__atomic_load_n(my_array+peer_indices[k], __ATOMIC_ACQUIRE)<value) {
			if(std::chrono::steady_clock::now()-start>std::chrono::milliseconds(timeout_ms>0 ? timeout_ms : 10000)) { my_array[63] = 1ull; break; }
			sched_yield();
		}
	}
	return FX3D_OK;
}
int fx3d_rendezvous_check(int, uint64_t* my_array, int) {
	if(my_array[63]!=0ull) { fx3d::set_error("halo rendezvous timed out: a neighbouring domain did not arrive"); return FX3D_ERR_TIMEOUT; }
	return FX3D_OK;
}

int fx3d_codec_fp16c_exhaustive(int, uint64_t*, uint32_t*) { fx3d::set_error("emulation: the exhaustive codec check runs on the GPU only"); return FX3D_ERR_INVALID; }

int fx3d_selftest_packed_math(int, uint64_t, uint64_t*) { fx3d::set_error("emulation: the packed-math self-test runs on the GPU only"); return FX3D_ERR_INVALID; }
int fx3d_selftest_division(int, uint64_t, uint64_t*) { fx3d::set_error("emulation: the division self-test runs on the GPU only"); return FX3D_ERR_INVALID; }

} // extern "C"

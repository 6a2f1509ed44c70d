// fx3d_runtime.cu -- devices, streams, events, memory, IPC and the device-side rendezvous of libfx3d_cuda.
// Replaces the reference's OpenCL wrapper: Device_Info / get_devices (FluidX3D v3.7 src/opencl.hpp:89-252), Device and
// its in-order queue (:284-340), Memory<T> allocation and transfers (:361-388,510-531,608-611), and the finish_queue
// barriers of the halo exchange (src/lbm.cpp:1357,1366,1375). CUDA only; there is no host fallback.
#define FX3D_TU_RUNTIME
#include <atomic>
#include "fx3d_internal.cuh"
#include <cstring>
#include <algorithm>

namespace fx3d {

int cuda_fail(cudaError_t e, const char* what) {
	(void)cudaGetLastError(); // a failed call also leaves its code in the runtime's last-error slot: clear it, or the next launch check reports it as its own
	set_error(std::string(what)+": "+cudaGetErrorName(e)+" ("+cudaGetErrorString(e)+")");
	if(e==cudaErrorNoDevice||e==cudaErrorInsufficientDriver||e==cudaErrorInvalidDevice) return FX3D_ERR_NO_DEVICE;
	if(e==cudaErrorMemoryAllocation) return FX3D_ERR_OUT_OF_MEMORY;
	return FX3D_ERR_CUDA;
}
#define FX3D_CUDA(call, what) do { const cudaError_t e_ = (call); if(e_!=cudaSuccess) return cuda_fail(e_, what); } while(0)

int use_device(int device) {
	FX3D_CUDA(cudaSetDevice(device), "cudaSetDevice");
	return FX3D_OK;
}
int check_launch(const char* what) {
	const cudaError_t e = cudaGetLastError();
	if(e!=cudaSuccess) return cuda_fail(e, what);
	return FX3D_OK;
}

__global__ void k_fill_f32(float* dst, float value, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x*blockDim.x;
	for(uint64_t k=(uint64_t)blockIdx.x*blockDim.x+threadIdx.x; k<n; k+=stride) dst[k] = value;
}

// rendezvous counters live in device memory that peers map (peer access or IPC); slot n_slots is the timeout flag
struct PeerList { uint64_t* p[32]; };
struct IndexList { int idx[32]; };
__global__ void k_rendezvous_signal(const PeerList peers, int n_peers, int my_index, uint64_t value) {
	const int k = (int)threadIdx.x;
	if(k<n_peers) {
		__threadfence_system(); // everything this stream wrote before is visible system-wide before the counter moves
		volatile uint64_t* slot = peers.p[k]+my_index;
		*slot = value;
		__threadfence_system();
	}
}
__global__ void k_rendezvous_wait(uint64_t* mine, const IndexList peer_indices, int n_peers, int error_slot, uint64_t value, long long timeout_cycles) {
	const int k = (int)threadIdx.x;
	if(k<n_peers) {
		volatile uint64_t* slot = mine+peer_indices.idx[k];
		const long long start = clock64();
		while(*slot<value) {
			if(clock64()-start>timeout_cycles) { mine[error_slot] = 1ull; break; }
			__nanosleep(200);
		}
		__threadfence_system();
	}
}

} // namespace fx3d
using namespace fx3d;

extern "C" {

int fx3d_device_count(int* count) {
	if(!count) return FX3D_ERR_INVALID;
	*count = 0;
	FX3D_CUDA(cudaGetDeviceCount(count), "cudaGetDeviceCount");
	if(*count<=0) { set_error("no CUDA device is available"); return FX3D_ERR_NO_DEVICE; }
	return FX3D_OK;
}
int fx3d_device_get_info(int device, fx3d_device_info* info) {
	if(!info) return FX3D_ERR_INVALID;
	cudaDeviceProp p;
	FX3D_CUDA(cudaGetDeviceProperties(&p, device), "cudaGetDeviceProperties");
	std::memset(info, 0, sizeof(*info));
	std::strncpy(info->name, p.name, sizeof(info->name)-1);
	info->id = device; info->cc_major = p.major; info->cc_minor = p.minor; info->sm_count = p.multiProcessorCount;
	static std::atomic<int> clock_khz[64]; // cudaDevAttrClockRate is not cached by the runtime (a query costs milliseconds): ask once per device
	int khz = device>=0&&device<64 ? clock_khz[device].load() : 0;
	if(khz==0) { cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device); if(device>=0&&device<64) clock_khz[device] = khz; }
	info->clock_mhz = khz/1000;
	info->memory_bytes = (uint64_t)p.totalGlobalMem; info->l2_bytes = (uint64_t)p.l2CacheSize;
	info->tflops_fp32 = (float)p.multiProcessorCount*128.0f*2.0f*(float)info->clock_mhz*1E-6f; // 128 FP32 lanes per SM, FMA = 2 flops
	return FX3D_OK;
}
int fx3d_device_enable_peer(int device, int peer) {
	if(device==peer) return FX3D_OK;
	int can = 0;
	FX3D_CUDA(cudaDeviceCanAccessPeer(&can, device, peer), "cudaDeviceCanAccessPeer");
	if(!can) { set_error("devices cannot access each other's memory"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(device)) return rc;
	const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
	if(e==cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return FX3D_OK; }
	if(e!=cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
	return FX3D_OK;
}
int fx3d_device_sync(int device) {
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
	return FX3D_OK;
}

int fx3d_stream_create(int device, fx3d_stream* stream) {
	if(!stream) return FX3D_ERR_INVALID;
	if(int rc = use_device(device)) return rc;
	cudaStream_t s;
	FX3D_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
	*stream = reinterpret_cast<fx3d_stream>(s);
	return FX3D_OK;
}
int fx3d_stream_destroy(int device, fx3d_stream stream) {
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaStreamDestroy(reinterpret_cast<cudaStream_t>(stream)), "cudaStreamDestroy");
	return FX3D_OK;
}
int fx3d_stream_sync(int device, fx3d_stream stream) {
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)), "cudaStreamSynchronize");
	return FX3D_OK;
}
int fx3d_event_create(int device, fx3d_event* event) {
	if(!event) return FX3D_ERR_INVALID;
	if(int rc = use_device(device)) return rc;
	cudaEvent_t e;
	FX3D_CUDA(cudaEventCreate(&e), "cudaEventCreate");
	*event = reinterpret_cast<fx3d_event>(e);
	return FX3D_OK;
}
int fx3d_event_destroy(int device, fx3d_event event) {
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event)), "cudaEventDestroy");
	return FX3D_OK;
}
int fx3d_event_record(int device, fx3d_event event, fx3d_stream stream) {
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), reinterpret_cast<cudaStream_t>(stream)), "cudaEventRecord");
	return FX3D_OK;
}
int fx3d_event_sync(int device, fx3d_event event) {
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)), "cudaEventSynchronize");
	return FX3D_OK;
}
int fx3d_event_elapsed_ms(fx3d_event start, fx3d_event stop, float* ms) {
	if(!ms) return FX3D_ERR_INVALID;
	FX3D_CUDA(cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(start), reinterpret_cast<cudaEvent_t>(stop)), "cudaEventElapsedTime");
	return FX3D_OK;
}
int fx3d_stream_wait_event(int device, fx3d_stream stream, fx3d_event event) {
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<cudaEvent_t>(event), 0), "cudaStreamWaitEvent");
	return FX3D_OK;
}

int fx3d_malloc(int device, size_t bytes, void** ptr) {
	if(!ptr) return FX3D_ERR_INVALID;
	*ptr = nullptr;
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaMalloc(ptr, bytes+256u), "cudaMalloc"); // 256 bytes of slack: bulk copies of rows that start off a 16-byte boundary (flag rows of x-decomposed domains) round their end up
	cudaError_t e = cudaMemset(*ptr, 0, bytes);
	if(e==cudaSuccess) e = cudaStreamSynchronize(0); // the library's streams are non-blocking: the zero fill must have landed before any of them touches the buffer
	if(e!=cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return cuda_fail(e, "cudaMemset"); }
	return FX3D_OK;
}
int fx3d_free(int device, void* ptr) {
	if(!ptr) return FX3D_OK;
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaFree(ptr), "cudaFree");
	return FX3D_OK;
}
int fx3d_host_alloc(size_t bytes, void** ptr) {
	if(!ptr) return FX3D_ERR_INVALID;
	*ptr = nullptr;
	FX3D_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1u, cudaHostAllocPortable), "cudaHostAlloc");
	return FX3D_OK;
}
int fx3d_host_free(void* ptr) {
	if(!ptr) return FX3D_OK;
	FX3D_CUDA(cudaFreeHost(ptr), "cudaFreeHost");
	return FX3D_OK;
}
static int copy(int device, void* dst, const void* src, size_t bytes, fx3d_stream stream, int blocking, cudaMemcpyKind kind) {
	if(bytes==0u) return FX3D_OK;
	if(!dst||!src) { set_error("memcpy with null pointer"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(device)) return rc;
	cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
	FX3D_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, s), "cudaMemcpyAsync");
	if(blocking) FX3D_CUDA(cudaStreamSynchronize(s), "cudaStreamSynchronize");
	return FX3D_OK;
}
int fx3d_memcpy_h2d(int device, void* dst, const void* src, size_t bytes, fx3d_stream stream, int blocking) { return copy(device, dst, src, bytes, stream, blocking, cudaMemcpyHostToDevice); }
int fx3d_memcpy_d2h(int device, void* dst, const void* src, size_t bytes, fx3d_stream stream, int blocking) { return copy(device, dst, src, bytes, stream, blocking, cudaMemcpyDeviceToHost); }
int fx3d_memset(int device, void* dst, int value, size_t bytes, fx3d_stream stream) {
	if(bytes==0u) return FX3D_OK;
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaMemsetAsync(dst, value, bytes, reinterpret_cast<cudaStream_t>(stream)), "cudaMemsetAsync");
	return FX3D_OK;
}
int fx3d_fill_f32(int device, float* dst, float value, size_t count, fx3d_stream stream) {
	if(count==0u) return FX3D_OK;
	if(int rc = use_device(device)) return rc;
	const uint32_t blocks = (uint32_t)std::min<uint64_t>((count+255u)/256u, 148u*16u);
	FX3D_LAUNCH(k_fill_f32, dim3(blocks), dim3(256u), stream, dst, value, (uint64_t)count);
	return check_launch("fill_f32");
}

int fx3d_ipc_get_handle(int device, void* ptr, void* handle64) {
	static_assert(sizeof(cudaIpcMemHandle_t)==64, "cudaIpcMemHandle_t is 64 bytes");
	if(!ptr||!handle64) return FX3D_ERR_INVALID;
	if(int rc = use_device(device)) return rc;
	cudaIpcMemHandle_t h;
	FX3D_CUDA(cudaIpcGetMemHandle(&h, ptr), "cudaIpcGetMemHandle");
	std::memcpy(handle64, &h, 64);
	return FX3D_OK;
}
int fx3d_ipc_open_handle(int device, const void* handle64, void** ptr) {
	if(!ptr||!handle64) return FX3D_ERR_INVALID;
	if(int rc = use_device(device)) return rc;
	cudaIpcMemHandle_t h;
	std::memcpy(&h, handle64, 64);
	FX3D_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
	return FX3D_OK;
}
int fx3d_ipc_close_handle(int device, void* ptr) {
	if(!ptr) return FX3D_OK;
	if(int rc = use_device(device)) return rc;
	FX3D_CUDA(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle");
	return FX3D_OK;
}

// rendezvous arrays hold 64 counters: slots 0..62 one per peer, slot 63 the timeout flag
int fx3d_rendezvous_signal(int device, uint64_t* const* peer_arrays, int n_peers, int my_index, uint64_t value, fx3d_stream stream) {
	if(n_peers<=0) return FX3D_OK;
	if(n_peers>32||!peer_arrays||my_index<0||my_index>62) { set_error("rendezvous: 1..32 peers, index 0..62"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(device)) return rc;
	PeerList pl;
	for(int k=0; k<32; k++) pl.p[k] = k<n_peers ? peer_arrays[k] : nullptr;
	FX3D_LAUNCH(k_rendezvous_signal, dim3(1u), dim3(32u), stream, pl, n_peers, my_index, value);
	return check_launch("rendezvous_signal");
}
int fx3d_rendezvous_wait(int device, uint64_t* my_array, const int* peer_indices, int n_peers, uint64_t value, int timeout_ms, fx3d_stream stream) {
	if(n_peers<=0) return FX3D_OK;
	if(n_peers>32||!my_array||!peer_indices) { set_error("rendezvous: 1..32 peers"); return FX3D_ERR_INVALID; }
	if(int rc = use_device(device)) return rc;
	IndexList il;
	for(int k=0; k<32; k++) il.idx[k] = k<n_peers ? peer_indices[k] : 0;
	static std::atomic<int> clock_khz[64]; // cudaDevAttrClockRate is not cached by the runtime (a query costs milliseconds): ask once per device
	int khz = device>=0&&device<64 ? clock_khz[device].load() : 0;
	if(khz==0) { cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device); if(device>=0&&device<64) clock_khz[device] = khz; }
	const long long cycles = (long long)(timeout_ms>0 ? timeout_ms : 10000)*(long long)(khz>0 ? khz : 1900000);
	FX3D_LAUNCH(k_rendezvous_wait, dim3(1u), dim3(32u), stream, my_array, il, n_peers, 63, value, cycles);
	return check_launch("rendezvous_wait");
}
int fx3d_rendezvous_check(int device, uint64_t* my_array, int n_slots) {
	(void)n_slots;
	if(int rc = use_device(device)) return rc;
	uint64_t flag = 0ull;
	FX3D_CUDA(cudaMemcpy(&flag, my_array+63, sizeof(flag), cudaMemcpyDeviceToHost), "cudaMemcpy(rendezvous flag)");
	if(flag!=0ull) { set_error("halo rendezvous timed out: a neighbouring domain did not arrive"); return FX3D_ERR_TIMEOUT; }
	return FX3D_OK;
}

int fx3d_codec_fp16c_exhaustive(int device, uint64_t* mismatches, uint32_t* first_bad_bits) {
	if(!mismatches||!first_bad_bits) return FX3D_ERR_INVALID;
	if(int rc = use_device(device)) return rc;
	unsigned long long* d_bad = nullptr; uint32_t* d_first = nullptr;
	FX3D_CUDA(cudaMalloc(&d_bad, sizeof(unsigned long long)), "cudaMalloc");
	FX3D_CUDA(cudaMalloc(&d_first, sizeof(uint32_t)), "cudaMalloc");
	cudaMemset(d_bad, 0, sizeof(unsigned long long)); cudaMemset(d_first, 0, sizeof(uint32_t));
	FX3D_LAUNCH(k_fp16c_exhaustive, dim3(148u*8u), dim3(256u), nullptr, d_bad, d_first);
	unsigned long long bad = 0ull;
	cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
	if(e==cudaSuccess) e = cudaMemcpy(first_bad_bits, d_first, sizeof(uint32_t), cudaMemcpyDeviceToHost);
	cudaFree(d_bad); cudaFree(d_first);
	if(e!=cudaSuccess) return cuda_fail(e, "fp16c exhaustive");
	*mismatches = (uint64_t)bad;
	return FX3D_OK;
}

int fx3d_selftest_packed_math(int device, uint64_t samples, uint64_t* mismatches) {
	if(!mismatches) return FX3D_ERR_INVALID;
	if(int rc = use_device(device)) return rc;
	unsigned long long* d_bad = nullptr;
	FX3D_CUDA(cudaMalloc(&d_bad, sizeof(unsigned long long)), "cudaMalloc");
	cudaMemset(d_bad, 0, sizeof(unsigned long long));
	const dim3 g(148u), b(128u);
	const unsigned long long per = (unsigned long long)(samples/(148u*128u)+1u);
	FX3D_LAUNCH((k_selftest_lanes<19, COLL_SRT, false>), g, b, nullptr, per, 1.0f, d_bad);
	FX3D_LAUNCH((k_selftest_lanes<19, COLL_SRT, false>), g, b, nullptr, per, 32768.0f, d_bad);
	FX3D_LAUNCH((k_selftest_lanes<19, COLL_TRT, true>), g, b, nullptr, per, 1.0f, d_bad);
	FX3D_LAUNCH((k_selftest_lanes<19, COLL_SRT, true>), g, b, nullptr, per, 32768.0f, d_bad);
	FX3D_LAUNCH((k_selftest_lanes<27, COLL_TRT, true>), g, b, nullptr, per, 32768.0f, d_bad);
	FX3D_LAUNCH((k_selftest_lanes<27, COLL_SRT, false>), g, b, nullptr, per, 1.0f, d_bad);
	FX3D_LAUNCH((k_selftest_lanes<27, COLL_TRT, false>), g, b, nullptr, per, 1.0f, d_bad);
	unsigned long long bad = 0ull;
	const cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
	cudaFree(d_bad);
	if(e!=cudaSuccess) return cuda_fail(e, "packed-math self-test");
	*mismatches = (uint64_t)bad;
	return FX3D_OK;
}
int fx3d_selftest_division(int device, uint64_t samples, uint64_t* mismatches) {
	if(!mismatches) return FX3D_ERR_INVALID;
	if(int rc = use_device(device)) return rc;
	unsigned long long* d_bad = nullptr;
	FX3D_CUDA(cudaMalloc(&d_bad, sizeof(unsigned long long)), "cudaMalloc");
	cudaMemset(d_bad, 0, sizeof(unsigned long long));
	const unsigned threads = 148u*8u*256u;
	FX3D_LAUNCH(k_selftest_division, dim3(148u*8u), dim3(256u), nullptr, (unsigned long long)((samples+threads-1u)/threads), d_bad);
	unsigned long long bad = 0ull;
	const cudaError_t e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
	cudaFree(d_bad);
	if(e!=cudaSuccess) return cuda_fail(e, "division self-test");
	*mismatches = (uint64_t)bad;
	return FX3D_OK;
}

} // extern "C"

// cuda.hpp -- device abstraction of the host surface: Device_Info, Device, Memory<T>. Drop-in for the reference's OpenCL
// wrapper (FluidX3D v3.7 src/opencl.hpp: Device_Info :89-198, get_devices :222-252, select_device_* :253-282,
// Device :284-340, Memory<T> :342-614) with the same member names, implemented on the C ABI of libfx3d_cuda.so
// (include/fx3d.h). Kernels are not looked up by name: LBM_Domain calls the typed fx3d_* launches directly.
// Error convention is the reference's: any failure prints a boxed message and exits (print_error).
#pragma once
#include "utilities.hpp"
#include "../../include/fx3d.h"

inline void fx3d_check(const int status, const string& what) {
	if(status!=FX3D_OK) print_error(what+": "+fx3d_last_error());
}

struct Device_Info {
	string name = "";
	uint id = 0u;             // CUDA ordinal
	uint memory = 0u;         // global memory in MB
	uint memory_used = 0u;    // MB allocated through Memory<T>
	uint compute_units = 0u;  // SMs
	uint clock_frequency = 0u; // MHz
	uint cores = 0u;          // FP32 lanes
	float tflops = 0.0f;
	bool is_gpu = true, is_cpu = false, uses_ram = false;
	uint max_workgroup_size = 1024u, local_cache = 227u, max_constant_buffer = 64u;
	string vendor = "NVIDIA";
	Device_Info() {}
	explicit Device_Info(const uint ordinal) {
		fx3d_device_info i;
		fx3d_check(fx3d_device_get_info((int)ordinal, &i), "device query");
		name = i.name; id = ordinal; memory = (uint)(i.memory_bytes/1048576ull); compute_units = (uint)i.sm_count;
		clock_frequency = (uint)i.clock_mhz; cores = compute_units*128u; tflops = i.tflops_fp32;
	}
};
inline vector<Device_Info> get_devices(const bool print_infos=false) {
	int count = 0;
	if(fx3d_device_count(&count)!=FX3D_OK) print_error("No CUDA devices are available. This build has no CPU path.");
	vector<Device_Info> devices;
	for(int d=0; d<count; d++) devices.push_back(Device_Info((uint)d));
	if(print_infos) for(const Device_Info& i : devices) print_info("Device ID "+to_string(i.id)+": "+i.name+" ("+to_string(i.memory)+" MB)");
	return devices;
}
inline Device_Info select_device_with_most_flops(const vector<Device_Info>& devices=get_devices()) {
	size_t best = 0u;
	for(size_t i=1u; i<devices.size(); i++) if(devices[i].tflops>devices[best].tflops) best = i;
	return devices[best];
}
inline Device_Info select_device_with_most_memory(const vector<Device_Info>& devices=get_devices()) {
	size_t best = 0u;
	for(size_t i=1u; i<devices.size(); i++) if(devices[i].memory>devices[best].memory) best = i;
	return devices[best];
}
inline Device_Info select_device_with_id(const uint id, const vector<Device_Info>& devices=get_devices()) {
	if(id>=(uint)devices.size()) print_error("Your selected Device ID ("+to_string(id)+") is wrong.");
	return devices[id];
}

class Device { // one in-order stream per device, like the reference's one queue per device
	fx3d_stream stream = nullptr;
	bool owns_stream = false;
public:
	Device_Info info;
	Device() {}
	explicit Device(const Device_Info& device_info, fx3d_stream shared_stream=nullptr) : info(device_info) {
		if(shared_stream) stream = shared_stream;
		else { fx3d_check(fx3d_stream_create((int)info.id, &stream), "stream creation"); owns_stream = true; }
	}
	fx3d_stream get_stream() const { return stream; }
	int ordinal() const { return (int)info.id; }
	void finish_queue() const { fx3d_check(fx3d_stream_sync((int)info.id, stream), "finish_queue"); }
	bool is_initialized() const { return stream!=nullptr; }
};

template<typename T> class Memory { // host (page-locked) + device buffer of N*dimensions elements of T
	ulong N = 0ull;
	uint d = 1u;
	bool host_buffer_exists = false, device_buffer_exists = false, external_host_buffer = false;
	T* host_buffer = nullptr;
	T* device_buffer = nullptr;
	Device* device = nullptr;
	void release() {
		if(device_buffer_exists && device_buffer) { fx3d_free(device->ordinal(), device_buffer); device->info.memory_used -= (uint)(capacity()/1048576ull); }
		if(host_buffer_exists && host_buffer && !external_host_buffer) fx3d_host_free(host_buffer);
		device_buffer = nullptr; host_buffer = nullptr; device_buffer_exists = host_buffer_exists = false;
	}
	void take(Memory& m) {
		N = m.N; d = m.d; host_buffer_exists = m.host_buffer_exists; device_buffer_exists = m.device_buffer_exists; external_host_buffer = m.external_host_buffer;
		host_buffer = m.host_buffer; device_buffer = m.device_buffer; device = m.device;
		m.host_buffer = nullptr; m.device_buffer = nullptr; m.host_buffer_exists = m.device_buffer_exists = false;
		set_pointers(); m.set_pointers();
	}
	void set_pointers() { // component planes of the host buffer (structure of arrays): x..w and s0..sF as in the reference's Memory<T> (src/opencl.hpp:356-359,392-393)
		T** const s[16] = { &s0, &s1, &s2, &s3, &s4, &s5, &s6, &s7, &s8, &s9, &sA, &sB, &sC, &sD, &sE, &sF };
		for(uint k=0u; k<16u; k++) *s[k] = (host_buffer && k<d) ? host_buffer+(ulong)k*N : nullptr;
		x = s0; y = s1; z = s2; w = s3;
	}
public:
	T *x=nullptr, *y=nullptr, *z=nullptr, *w=nullptr; // host pointers to the component planes (SoA)
	T *s0=nullptr, *s1=nullptr, *s2=nullptr, *s3=nullptr, *s4=nullptr, *s5=nullptr, *s6=nullptr, *s7=nullptr, *s8=nullptr, *s9=nullptr, *sA=nullptr, *sB=nullptr, *sC=nullptr, *sD=nullptr, *sE=nullptr, *sF=nullptr;
	Memory() {}
	Memory(Device& dev, const ulong N_, const uint dimensions=1u, const bool allocate_host=true, const bool allocate_device=true, const T value=(T)0) : N(N_), d(dimensions), device(&dev) {
		if(N*(ulong)d==0ull) print_error("Memory size must be larger than 0.");
		if(allocate_device) {
			void* p = nullptr;
			const int rc = fx3d_malloc(dev.ordinal(), (size_t)capacity(), &p);
			if(rc==FX3D_ERR_OUT_OF_MEMORY) print_error("Memory size is too large at "+to_string((uint)(capacity()/1048576ull))+" MB. Device \""+dev.info.name+"\" does not have enough memory. Allocating another "+to_string((uint)(capacity()/1048576ull))+" MB would use a total of "+to_string(dev.info.memory_used+(uint)(capacity()/1048576ull))+" MB / "+to_string(dev.info.memory)+" MB.");
			fx3d_check(rc, "device memory allocation");
			device_buffer = (T*)p; device_buffer_exists = true;
			dev.info.memory_used += (uint)(capacity()/1048576ull);
		}
		if(allocate_host) {
			void* p = nullptr;
			fx3d_check(fx3d_host_alloc((size_t)capacity(), &p), "host memory allocation");
			host_buffer = (T*)p; host_buffer_exists = true;
			for(ulong i=0ull; i<range(); i++) host_buffer[i] = value;
			set_pointers();
		}
		if(allocate_device && value!=(T)0) { // device buffers come zero-filled from fx3d_malloc
			if(host_buffer_exists) write_to_device();
			else if(sizeof(T)==4u) { float f; std::memcpy(&f, &value, 4); fx3d_check(fx3d_fill_f32(dev.ordinal(), (float*)device_buffer, f, (size_t)range(), dev.get_stream()), "fill"); }
		}
	}
	~Memory() { release(); }
	Memory(const Memory&) = delete;
	Memory& operator=(const Memory&) = delete;
	Memory(Memory&& m) noexcept { take(m); }
	Memory& operator=(Memory&& m) noexcept { if(this!=&m) { release(); take(m); } return *this; }
	ulong length() const { return N; }
	uint dimensions() const { return d; }
	ulong range() const { return N*(ulong)d; }
	ulong capacity() const { return N*(ulong)d*sizeof(T); }
	T* data() { return host_buffer; }
	const T* data() const { return host_buffer; }
	T* device_data() const { return device_buffer; } // device pointer, handed to the fx3d_lattice of the owning domain
	T& operator[](const ulong i) { return host_buffer[i]; }
	const T& operator[](const ulong i) const { return host_buffer[i]; }
	T* operator()() { return host_buffer; } // (src/opencl.hpp:504-509)
	const T* operator()() const { return host_buffer; }
	T operator()(const ulong i) const { return host_buffer[i]; }
	T operator()(const ulong i, const uint dimension) const { return host_buffer[i+(ulong)dimension*N]; } // component `dimension` of element i
	T* exchange_host_buffer(T* other) { T* mine = host_buffer; host_buffer = other; external_host_buffer = true; set_pointers(); return mine; }
	void reset(const T value=(T)0) {
		if(host_buffer_exists) for(ulong i=0ull; i<range(); i++) host_buffer[i] = value;
		if(device_buffer_exists) { if(host_buffer_exists) write_to_device(); else fx3d_check(fx3d_memset(device->ordinal(), device_buffer, 0, (size_t)capacity(), device->get_stream()), "reset"); }
	}
	void enqueue_read_from_device(const ulong offset, const ulong length) {
		if(host_buffer_exists&&device_buffer_exists) fx3d_check(fx3d_memcpy_d2h(device->ordinal(), host_buffer+offset, device_buffer+offset, (size_t)(min(length, range()-offset)*sizeof(T)), device->get_stream(), 0), "read_from_device");
	}
	void enqueue_write_to_device(const ulong offset, const ulong length) {
		if(host_buffer_exists&&device_buffer_exists) fx3d_check(fx3d_memcpy_h2d(device->ordinal(), device_buffer+offset, host_buffer+offset, (size_t)(min(length, range()-offset)*sizeof(T)), device->get_stream(), 0), "write_to_device");
	}
	void enqueue_read_from_device() { enqueue_read_from_device(0ull, range()); }
	void enqueue_write_to_device() { enqueue_write_to_device(0ull, range()); }
	void read_from_device(const bool blocking=true) { enqueue_read_from_device(); if(blocking) finish_queue(); }
	void write_to_device(const bool blocking=true) { enqueue_write_to_device(); if(blocking) finish_queue(); }
	void read_from_device(const ulong offset, const ulong length, const bool blocking=true) { enqueue_read_from_device(offset, length); if(blocking) finish_queue(); }
	void write_to_device(const ulong offset, const ulong length, const bool blocking=true) { enqueue_write_to_device(offset, length); if(blocking) finish_queue(); }
	void finish_queue() { device->finish_queue(); }
};

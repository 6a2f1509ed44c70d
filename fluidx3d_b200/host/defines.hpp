// defines.hpp -- compile-time configuration, same switch names as the reference (FluidX3D v3.7 src/defines.hpp) so that a
// scene written for it builds here unchanged. Switches outside the B200 hot path still exist, and stop the build with a
// clear message when enabled.
#pragma once

#ifndef FX3D_CUSTOM_DEFINES // build with -DFX3D_CUSTOM_DEFINES -DD3Q27 -DTRT ... to choose the switches on the command line instead

// velocity set (exactly one)
#define D3Q19
//#define D3Q27

// collision operator (exactly one)
#define SRT
//#define TRT

// DDF storage: none = FP32, or one of the two 16-bit compressions (arithmetic stays FP32)
#define FP16S
//#define FP16C

#define BENCHMARK // run the benchmark scene with all extensions off

// extensions on the hot path
//#define VOLUME_FORCE
//#define EQUILIBRIUM_BOUNDARIES
//#define UPDATE_FIELDS
//#define SUBGRID // Smagorinsky-Lilly subgrid turbulence model (runs the whole-row bulk-copy kernel or the general kernel)
//#define MOVING_BOUNDARIES // moving solid boundaries (TYPE_S cells with non-zero velocity)
//#define FORCE_FIELD // boundary forces on solid cells (lbm.object_force() etc.); with VOLUME_FORCE also a per-cell force lbm.F

// extensions of the reference that this build does not provide
//#define SURFACE
//#define TEMPERATURE
//#define PARTICLES
//#define INTERACTIVE_GRAPHICS
//#define INTERACTIVE_GRAPHICS_ASCII
//#define GRAPHICS

#endif // FX3D_CUSTOM_DEFINES

// ---------------------------------------------------------------------------------------------------------------------

#define TYPE_S 0b00000001 // solid boundary
#define TYPE_E 0b00000010 // equilibrium boundary (inflow/outflow)
#define TYPE_T 0b00000100 // temperature boundary (TEMPERATURE only)
#define TYPE_F 0b00001000 // fluid   (SURFACE only)
#define TYPE_I 0b00010000 // interface (SURFACE only)
#define TYPE_G 0b00100000 // gas     (SURFACE only)
#define TYPE_X 0b01000000 // user marker X
#define TYPE_Y 0b10000000 // user marker Y

#if defined(FP16S) || defined(FP16C)
#define fpxx ushort
#else
#define fpxx float
#endif

#ifdef BENCHMARK
#undef UPDATE_FIELDS
#undef VOLUME_FORCE
#undef EQUILIBRIUM_BOUNDARIES
#undef FORCE_FIELD
#undef MOVING_BOUNDARIES
#undef SURFACE
#undef TEMPERATURE
#undef SUBGRID
#undef PARTICLES
#undef INTERACTIVE_GRAPHICS
#undef INTERACTIVE_GRAPHICS_ASCII
#undef GRAPHICS
#endif

#if defined(SURFACE) || defined(TEMPERATURE) || defined(PARTICLES)
#error "SURFACE / TEMPERATURE / PARTICLES are not part of the B200 hot-path build (see DESIGN.md, out of scope)"
#endif
#if defined(INTERACTIVE_GRAPHICS) || defined(INTERACTIVE_GRAPHICS_ASCII) || defined(GRAPHICS)
#error "graphics are not part of the B200 hot-path build (see DESIGN.md, out of scope)"
#endif
#if defined(D2Q9) || defined(D3Q15)
#error "only D3Q19 and D3Q27 are on the B200 hot path"
#endif
#if defined(D3Q19) == defined(D3Q27)
#error "select exactly one of D3Q19 / D3Q27"
#endif
#if defined(FP16S) && defined(FP16C)
#error "select at most one of FP16S / FP16C"
#endif

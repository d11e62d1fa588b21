// fgs_launch.h -- kernel launch and error reporting shared by the self-contained operator files (loss.cu,
// densify.cu).  These files are also compiled by g++ against tests/cuda_emu/cuda_emu.h (a test-only
// thread-per-CUDA-thread emulator, FGS_CUDA_EMU defined) so that their indexing and launch logic can be checked
// without a GPU; the emulator provides its own FGS_LAUNCH.
#pragma once
#ifndef FGS_CUDA_EMU
#include <cuda_runtime.h>
#define FGS_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
// same launch; tells the emulator that the kernel has no barrier and no warp collective, so it may run the CUDA
// threads one after the other instead of as host threads
#define FGS_LAUNCH_SERIAL(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
// launch with `smem` bytes of dynamic shared memory, declared inside the kernel with FGS_DYNAMIC_SMEM
#define FGS_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FGS_DYNAMIC_SMEM(type, name, alignment) extern __shared__ __align__(alignment) type name[]
#endif

namespace fgs {
// records the text returned by fgs_last_error() and returns `code` (c_abi.cu; the emulator harness has its own)
int report_error(int code, const char* where, const char* detail);
}  // namespace fgs

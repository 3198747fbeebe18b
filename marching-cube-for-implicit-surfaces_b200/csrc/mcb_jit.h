/* Run-time specialisation of the field evaluator: the fused grid program of ONE equation as a
 * straight-line sm_100a kernel, compiled with NVRTC — same tile, same operand fetches, same fp32 operations in the
 * same order as the interpreter (eval_field_kernel), minus the dispatch.  Host side only; mcb_api.cu loads the cubin. */
#ifndef MCB_JIT_H
#define MCB_JIT_H

#include <string>
#include <vector>

#include "mcb_bytecode.h"

namespace mcbjit {

/* CUDA source of the specialised kernels (`mcb_eval_jit`, `mcb_signs_jit`, `mcb_fill_jit`) for a fused grid program (words as mcb::fuse() leaves them:
 * fop | src << 4 | arg << 8, table operands still as (axis, slot)).  The source depends on the program only — constants,
 * grid size and table offsets are kernel arguments — so an equation is compiled once.  Empty string + *err on a
 * program the generator does not take (raw coordinate operands). */
std::string generate(const uint32_t* code, int n, bool* has_pow, std::string* err);

/* NVRTC (libnvrtc.so.12, loaded on first use) -> cubin for sm_100a.  Returns "" on success, else the error / compile log. */
std::string compile(const std::string& source, bool has_pow, int grid_bytes, std::vector<char>* cubin);

} /* namespace mcbjit */
#endif

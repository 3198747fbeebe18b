/* Bytecode of the field evaluator: what the host lowering (mcb_lower.cpp) emits and the device interprets.
 *
 * One instruction = one 32-bit word: opcode in bits 0..7, argument in bits 8..31.  A program is a postfix
 * sequence over an operand stack; a binary opcode pops `top` then `second` and pushes
 *      second (op) top         for OP_ADD/SUB/MUL/DIV/POW
 *      top (op) second         for the reversed forms OP_RSUB/RDIV/RPOW
 * so the lowering can evaluate the operands of the reference's `val2 (op) val1` (evaluator.cpp:38-40) in either
 * order while each value keeps its role.  All arithmetic is fp32 without FMA contraction; `^` is mcb_powf.
 *
 * Programs travel to the kernels inside the kernel-parameter block, i.e. in the constant bank: every lane of a
 * warp executes the same instruction word, fetched with a uniform constant load.
 */
#ifndef MCB_BYTECODE_H
#define MCB_BYTECODE_H

#include <stdint.h>

#include "mcb_pow.h"

enum {
    MCB_OP_END = 0,
    MCB_OP_PUSH_X,  /* scaled coordinate sx*x (marching.cpp:211) */
    MCB_OP_PUSH_Y,
    MCB_OP_PUSH_Z,
    MCB_OP_PUSH_K,  /* arg = constant-pool index (literal or folded constant subtree) */
    MCB_OP_PUSH_TX, /* arg = axis-table slot: value of a hoisted x-only subtree at this vertex's x index */
    MCB_OP_PUSH_TY,
    MCB_OP_PUSH_TZ,
    MCB_OP_ADD,
    MCB_OP_SUB,
    MCB_OP_RSUB,
    MCB_OP_MUL,
    MCB_OP_DIV,
    MCB_OP_RDIV,
    MCB_OP_POW,
    MCB_OP_RPOW,
    MCB_OP_NEG,
    MCB_OP_COUNT
};

#define MCB_MAX_CODE 384   /* instructions per program (a 256-char equation yields < 300) */
#define MCB_MAX_K 128      /* constant pool entries */
#define MCB_MAX_SLOTS 96   /* hoisted subtrees (constant + per-axis) */
#define MCB_MAX_STACK 24   /* operand stack depth any program may need */

#define MCB_INSN(op, arg) ((uint32_t)(op) | ((uint32_t)(arg) << 8))
#define MCB_INSN_OP(w) ((w) & 0xFFu)
#define MCB_INSN_ARG(w) ((w) >> 8)

/* A program + constant pool, passed by value as a __grid_constant__ kernel parameter (constant bank). */
typedef struct {
    int n;
    uint32_t code[MCB_MAX_CODE];
    float k[MCB_MAX_K];
} mcb_program;

#if defined(__CUDACC__)
#define MCB_BC_FN __host__ __device__ __forceinline__
#else
#define MCB_BC_FN static inline
#endif

MCB_BC_FN float mcb_binop(uint32_t op, float second, float top) {
    switch (op) {
        case MCB_OP_ADD: return second + top;
        case MCB_OP_SUB: return second - top;
        case MCB_OP_RSUB: return top - second;
        case MCB_OP_MUL: return second * top;
        case MCB_OP_DIV: return second / top;
        case MCB_OP_RDIV: return top / second;
        case MCB_OP_POW: return mcb_powf(second, top);
        default: return mcb_powf(top, second); /* MCB_OP_RPOW */
    }
}

/* Plain scalar interpreter: one point, memory stack.  Used for the rare, divergent evaluations (ambiguity face
 * centres, mcb_eval_points, constant folding, axis tables); the per-vertex grid kernel has its own register-cached
 * version.  tx/ty/tz = this point's table values per slot (may be NULL when the program has no table loads). */
MCB_BC_FN float mcb_interp_scalar(const uint32_t* code, int n, const float* k, float x, float y, float z,
                                  const float* tx, const float* ty, const float* tz) {
    float st[MCB_MAX_STACK];
    int sp = 0;
    for (int pc = 0; pc < n; pc++) {
        uint32_t w = code[pc], op = MCB_INSN_OP(w), arg = MCB_INSN_ARG(w);
        switch (op) {
            case MCB_OP_PUSH_X: st[sp++] = x; break;
            case MCB_OP_PUSH_Y: st[sp++] = y; break;
            case MCB_OP_PUSH_Z: st[sp++] = z; break;
            case MCB_OP_PUSH_K: st[sp++] = k[arg]; break;
            case MCB_OP_PUSH_TX: st[sp++] = tx[arg]; break;
            case MCB_OP_PUSH_TY: st[sp++] = ty[arg]; break;
            case MCB_OP_PUSH_TZ: st[sp++] = tz[arg]; break;
            case MCB_OP_NEG: st[sp - 1] = -st[sp - 1]; break;
            case MCB_OP_END: pc = n; break;
            default: {
                float top = st[--sp];
                st[sp - 1] = mcb_binop(op, st[sp - 1], top);
            }
        }
    }
    return st[0];
}

#endif /* MCB_BYTECODE_H */

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

/* ---- fused ("accumulator") form of a grid program -----------------------------------------------------------
 * The grid kernel does not run the postfix words above directly.  mcb::fuse() (mcb_lower.cpp) rewrites them so
 * that a push which is immediately consumed by a binary operator becomes that operator's second operand:
 *      word = fop | src << 4 | arg << 8
 * The machine state is an accumulator `acc` (= the top of the operand stack, in registers) and a memory stack
 * holding the deeper levels.  `v` is the operand named by src:
 *      src X,Y,Z     scaled coordinate of the vertex            src K       constant pool entry arg
 *      src TX,TY,TZ  axis table arg at the vertex's index       src POP     pop the memory stack
 *      fop LOAD      acc = v                                    fop PUSH    spill acc to the memory stack; acc = v
 *      fop ADD, MUL  acc = acc (op) v                           fop NEG     acc = -acc  (src unused)
 *      fop SUB/DIV/POW  acc = acc (op) v                        fop RSUB/RDIV/RPOW  acc = v (op) acc
 * Every fp32 operation and the role of each operand are those of the postfix program; only pushes and pops of
 * the operand stack disappear.  `x^2+y^2+z^2-0.49` (grid form TX0 TY0 TZ0 K0 - + +) becomes
 *      LOAD TZ0; SUB K0; ADD TY0; ADD TX0 — 4 words and no memory-stack traffic at all.
 */
enum { MCB_SRC_X = 0, MCB_SRC_Y, MCB_SRC_Z, MCB_SRC_K, MCB_SRC_TX, MCB_SRC_TY, MCB_SRC_TZ, MCB_SRC_POP };
enum {
    MCB_F_LOAD = 0, MCB_F_PUSH, MCB_F_ADD, MCB_F_SUB, MCB_F_RSUB, MCB_F_MUL, MCB_F_DIV, MCB_F_RDIV, MCB_F_POW,
    MCB_F_RPOW, MCB_F_NEG, MCB_F_COUNT
};
#define MCB_FINSN(fop, src, arg) ((uint32_t)(fop) | ((uint32_t)(src) << 4) | ((uint32_t)(arg) << 8))
#define MCB_FINSN_OP(w) ((w) & 0xFu)
#define MCB_FINSN_SRC(w) (((w) >> 4) & 0xFu)
#define MCB_FINSN_ARG(w) ((w) >> 8)

MCB_BC_FN float mcb_fop(uint32_t fop, float acc, float v) {
    switch (fop) {
        case MCB_F_ADD: return acc + v;
        case MCB_F_SUB: return acc - v;
        case MCB_F_RSUB: return v - acc;
        case MCB_F_MUL: return acc * v;
        case MCB_F_DIV: return acc / v;
        case MCB_F_RDIV: return v / acc;
        case MCB_F_POW: return mcb_powf(acc, v);
        case MCB_F_RPOW: return mcb_powf(v, acc);
        default: return v; /* LOAD, PUSH */
    }
}

/* Scalar interpreter of the fused form (host-side checks of mcb::fuse(); the grid kernel has its own). */
MCB_BC_FN float mcb_interp_fused_scalar(const uint32_t* code, int n, const float* k, float x, float y, float z,
                                        const float* tx, const float* ty, const float* tz) {
    float st[MCB_MAX_STACK];
    int sp = 0;
    float acc = 0.f;
    for (int pc = 0; pc < n; pc++) {
        uint32_t w = code[pc], fop = MCB_FINSN_OP(w), src = MCB_FINSN_SRC(w), arg = MCB_FINSN_ARG(w);
        if (fop == MCB_F_NEG) { acc = -acc; continue; }
        if (fop == MCB_F_PUSH) st[sp++] = acc;
        float v;
        switch (src) {
            case MCB_SRC_X: v = x; break;
            case MCB_SRC_Y: v = y; break;
            case MCB_SRC_Z: v = z; break;
            case MCB_SRC_K: v = k[arg]; break;
            case MCB_SRC_TX: v = tx[arg]; break;
            case MCB_SRC_TY: v = ty[arg]; break;
            case MCB_SRC_TZ: v = tz[arg]; break;
            default: v = st[--sp]; break;
        }
        acc = mcb_fop(fop, acc, v);
    }
    return acc;
}

#endif /* MCB_BYTECODE_H */

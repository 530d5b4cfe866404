/* NVRTC stand-in for include/llo_cuda.h: the constants the device code uses.
 * (the real header declares host functions, which a JIT translation unit may
 * not contain; tests/test_capi_symbols.py checks the two stay in step) */
#ifndef LLO_CUDA_H
#define LLO_CUDA_H

#include <stdint.h>

#define LLO_RANK 8
#define LLO_MAT 81

enum {
	LLO_OP_BAD = 0, LLO_OP_ABS, LLO_OP_NEG, LLO_OP_SIN, LLO_OP_COS, LLO_OP_TAN,
	LLO_OP_EXP, LLO_OP_LOG, LLO_OP_SQRT, LLO_OP_ROUND, LLO_OP_POW, LLO_OP_SUM,
	LLO_OP_SUB, LLO_OP_PROD, LLO_OP_DIV, LLO_OP_MIN, LLO_OP_MAX, LLO_OP_EQ,
	LLO_OP_NEQ, LLO_OP_LT, LLO_OP_GT, LLO_OP_RAND_BINO, LLO_OP_RAND_UNIF,
	LLO_OP_RAND_NORM, LLO_OP_COUNT
};

typedef struct llo_cuda_insn { uint8_t op; uint8_t arg; } llo_cuda_insn;
#define LLO_VM_MAX_INPUTS 8
#define LLO_VM_MAX_INSNS 48
#define LLO_VM_MAX_CONSTS 8
#define LLO_VM_MAX_DEPTH 4

#endif

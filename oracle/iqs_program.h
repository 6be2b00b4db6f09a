/*
 * iqs_program.h -- flat description of a circuit ("program") replayed by
 *   - oracle/iqs_oracle.c   (the CPU restatement of the reference algorithm),
 *   - oracle/driver.cpp     (the same source compiled against the REAL reference library
 *                            -> oracle/_ref/iqs_ref_driver, and against the B200 drop-in
 *                            library -> intel-qs_b200/bin/iqs_b200_driver),
 *   - the Python harness    (tests/, bench.py) through the C ABI.
 * TEST INFRASTRUCTURE ONLY: nothing in the product path includes this file.
 *
 * Every op names PROGRAM qubits, exactly like the public iqs::QubitRegister methods
 * (reference include/qureg.hpp:212-331).
 */
#ifndef IQS_PROGRAM_H
#define IQS_PROGRAM_H
#include <stdint.h>

typedef struct iqs_op {
  int32_t kind;
  int32_t q0, q1, q2;
  double p[40]; /* matrix (row-major, re/im interleaved), angles, or a qubit map (<= 40 qubits) */
} iqs_op;

enum {
  /* generic gates */
  OP_GATE1 = 1,    /* Apply1QubitGate(q0, m = p[0..7])                       */
  OP_CGATE1 = 2,   /* ApplyControlled1QubitGate(control q0, target q1, m)    */
  OP_SWAPLIKE = 3, /* ApplySwap_helper(q0, q1, m)                            */
  OP_DIAG = 4,     /* ApplyDiag(q0, q1, diag(p[0..7]))                       */
  OP_GATE2 = 5,    /* Apply2QubitGate(q0, q1, m = p[0..31])                  */
  OP_TOFFOLI = 6,  /* ApplyToffoli(q0, q1, q2)                               */
  /* named 1-qubit gates (reference src/qureg_apply1qubitgate.cpp:277-501) */
  OP_H = 10, OP_X = 11, OP_Y = 12, OP_Z = 13, OP_SQRTX = 14, OP_SQRTY = 15, OP_SQRTZ = 16, OP_T = 17,
  OP_RX = 18, OP_RY = 19, OP_RZ = 20, /* angle p[0] */
  OP_RXY = 21,                        /* phi p[0], theta p[1] */
  /* named controlled gates (src/qureg_applyctrl1qubitgate.cpp:458-618): control q0, target q1 */
  OP_CH = 30, OP_CX = 31, OP_CY = 32, OP_CZ = 33, OP_CSQRTZ = 34,
  OP_CRX = 35, OP_CRY = 36, OP_CRZ = 37, OP_CPHASE = 38, /* angle p[0] */
  /* swap family (src/qureg_applyswap.cpp:23-108) */
  OP_SWAP = 40, OP_ISWAP = 41, OP_SQRTISWAP = 42, OP_4THROOTISWAP = 43,
  /* value-returning: each appends one double to the scalar output */
  OP_PROB = 50,      /* GetProbability(q0)                                   */
  OP_EXPECT = 51,    /* ExpectationValue(qubits p[0..q0-1], observables p[16..16+q0-1], coeff 1) */
  OP_NORM = 52,      /* ComputeNorm()                                        */
  OP_NORMALIZE = 53, /* Normalize()                                          */
  OP_COLLAPSE = 54,  /* CollapseQubit(q0, q1 != 0)                           */
  OP_EXPECT1 = 55,   /* ExpectationValueX/Y/Z(q0), observable q1 in 1..3     */
  OP_ENTROPY = 56,     /* Entropy(): one scalar (src/qureg_utils.cpp:305-345)  */
  OP_GOOGLESTATS = 57, /* GoogleStats(): eleven scalars (:350-450)             */
  OP_GETAMP = 58,      /* GetGlobalAmplitude(index = p[0], exact below 2^53): two scalars (re, im) */
  /* qubit order */
  OP_PERMUTE = 60,  /* PermuteQubits(map = p[0..n-1], "direct")              */
  OP_EMUSWAP = 61,  /* EmulateSwap(q0, q1)                                   */
  /* modes */
  OP_FUSION_ON = 70, /* TurnOnFusion(q0)  */
  OP_FUSION_OFF = 71,
  OP_SPEC_ON = 72, OP_SPEC_OFF = 73, OP_SPEC2_ON = 74, OP_SPEC2_OFF = 75
};

/* file layout of a program: header then nops ops */
typedef struct iqs_program_header {
  uint32_t magic; /* 'IQSP' = 0x50535149 */
  uint32_t num_qubits;
  uint32_t nops;
  uint32_t init; /* 0: state read from the state file; 1: "base" with base_index; 2: "++++" */
  uint64_t base_index;
  uint64_t reserved;
} iqs_program_header;
#define IQS_PROGRAM_MAGIC 0x50535149u

#endif

// driver.cpp -- replay a program file (oracle/iqs_program.h) through the public
// iqs::QubitRegister<ComplexDP> API and dump the resulting state, qubit map and scalars.
//
// TEST INFRASTRUCTURE.  The SAME source is compiled twice:
//   * against the unmodified reference (headers + sources under /root/reference)
//       -> oracle/_ref/iqs_ref_driver      (pins the oracle, generates tests/golden, CPU baseline)
//   * against the B200 drop-in (intel-qs_b200/include + libiqs.so over libiqs_b200.so)
//       -> intel-qs_b200/bin/iqs_b200_driver (the drop-in proof: no source change)
//
// usage: driver <program.bin> [--state-in s.bin] [--state-out s.bin] [--scalars-out x.bin]
//               [--map-out m.bin] [--repeat R] [--step-sizes a,b,c] [--no-step-norm]
// Prints one line "TIME <seconds> OPS <count>" for the timed replay (init excluded).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "qureg.hpp"

#include "iqs_program.h"

using Reg = iqs::QubitRegister<ComplexDP>;

static TM2x2<ComplexDP> mat2(const double *p) {
  TM2x2<ComplexDP> m;
  m(0, 0) = ComplexDP(p[0], p[1]);
  m(0, 1) = ComplexDP(p[2], p[3]);
  m(1, 0) = ComplexDP(p[4], p[5]);
  m(1, 1) = ComplexDP(p[6], p[7]);
  return m;
}

static void run_ops(Reg &psi, const std::vector<iqs_op> &ops, std::vector<double> &scalars) {
  const unsigned n = (unsigned)psi.NumQubits();
  for (const iqs_op &op : ops) {
    switch (op.kind) {
      case OP_GATE1: psi.Apply1QubitGate(op.q0, mat2(op.p)); break;
      case OP_CGATE1: psi.ApplyControlled1QubitGate(op.q0, op.q1, mat2(op.p)); break;
      case OP_SWAPLIKE: psi.ApplySwap_helper(op.q0, op.q1, mat2(op.p)); break;
      case OP_DIAG: {
        TM4x4<ComplexDP> d;
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j) d(i, j) = ComplexDP(0, 0);
        for (int i = 0; i < 4; ++i) d(i, i) = ComplexDP(op.p[2 * i], op.p[2 * i + 1]);
        psi.ApplyDiag(op.q0, op.q1, d);
        break;
      }
      case OP_GATE2: {
        TM4x4<ComplexDP> m;
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j) m(i, j) = ComplexDP(op.p[2 * (4 * i + j)], op.p[2 * (4 * i + j) + 1]);
        psi.Apply2QubitGate(op.q0, op.q1, m);
        break;
      }
      case OP_TOFFOLI: psi.ApplyToffoli(op.q0, op.q1, op.q2); break;
      case OP_H: psi.ApplyHadamard(op.q0); break;
      case OP_X: psi.ApplyPauliX(op.q0); break;
      case OP_Y: psi.ApplyPauliY(op.q0); break;
      case OP_Z: psi.ApplyPauliZ(op.q0); break;
      case OP_SQRTX: psi.ApplyPauliSqrtX(op.q0); break;
      case OP_SQRTY: psi.ApplyPauliSqrtY(op.q0); break;
      case OP_SQRTZ: psi.ApplyPauliSqrtZ(op.q0); break;
      case OP_T: psi.ApplyT(op.q0); break;
      case OP_RX: psi.ApplyRotationX(op.q0, op.p[0]); break;
      case OP_RY: psi.ApplyRotationY(op.q0, op.p[0]); break;
      case OP_RZ: psi.ApplyRotationZ(op.q0, op.p[0]); break;
      case OP_RXY: psi.ApplyRotationXY(op.q0, op.p[0], op.p[1]); break;
      case OP_CH: psi.ApplyCHadamard(op.q0, op.q1); break;
      case OP_CX: psi.ApplyCPauliX(op.q0, op.q1); break;
      case OP_CY: psi.ApplyCPauliY(op.q0, op.q1); break;
      case OP_CZ: psi.ApplyCPauliZ(op.q0, op.q1); break;
      case OP_CSQRTZ: psi.ApplyCPauliSqrtZ(op.q0, op.q1); break;
      case OP_CRX: psi.ApplyCRotationX(op.q0, op.q1, op.p[0]); break;
      case OP_CRY: psi.ApplyCRotationY(op.q0, op.q1, op.p[0]); break;
      case OP_CRZ: psi.ApplyCRotationZ(op.q0, op.q1, op.p[0]); break;
      case OP_CPHASE: psi.ApplyCPhaseRotation(op.q0, op.q1, op.p[0]); break;
      case OP_SWAP: psi.ApplySwap(op.q0, op.q1); break;
      case OP_ISWAP: psi.ApplyISwap(op.q0, op.q1); break;
      case OP_SQRTISWAP: psi.ApplySqrtISwap(op.q0, op.q1); break;
      case OP_4THROOTISWAP: psi.Apply4thRootISwap(op.q0, op.q1); break;
      case OP_PROB: scalars.push_back(psi.GetProbability(op.q0)); break;
      case OP_EXPECT: {
        std::vector<unsigned> qs, obs;
        for (int i = 0; i < op.q0; ++i) {
          qs.push_back((unsigned)op.p[i]);
          obs.push_back((unsigned)op.p[16 + i]);
        }
        scalars.push_back(psi.ExpectationValue(qs, obs, 1.));
        break;
      }
      case OP_EXPECT1:
        if (op.q1 == 1) scalars.push_back(psi.ExpectationValueX(op.q0, 1.));
        else if (op.q1 == 2) scalars.push_back(psi.ExpectationValueY(op.q0, 1.));
        else scalars.push_back(psi.ExpectationValueZ(op.q0, 1.));
        break;
      case OP_NORM: scalars.push_back(psi.ComputeNorm()); break;
      case OP_ENTROPY: scalars.push_back(psi.Entropy()); break;
      case OP_GOOGLESTATS: {
        std::vector<double> st = psi.GoogleStats();
        scalars.insert(scalars.end(), st.begin(), st.end());
        break;
      }
      case OP_GETAMP: {
        ComplexDP a = psi.GetGlobalAmplitude((std::size_t)op.p[0]);
        scalars.push_back(a.real());
        scalars.push_back(a.imag());
        break;
      }
      case OP_NORMALIZE: psi.Normalize(); break;
      case OP_COLLAPSE: psi.CollapseQubit(op.q0, op.q1 != 0); break;
      case OP_PERMUTE: {
        std::vector<std::size_t> map(n);
        for (unsigned i = 0; i < n; ++i) map[i] = (std::size_t)op.p[i];
        psi.PermuteQubits(map, "direct");
        break;
      }
      case OP_EMUSWAP: psi.EmulateSwap(op.q0, op.q1); break;
      case OP_FUSION_ON: psi.TurnOnFusion(op.q0); break;
      case OP_FUSION_OFF: psi.TurnOffFusion(); break;
      case OP_SPEC_ON: psi.TurnOnSpecialize(); break;
      case OP_SPEC_OFF: psi.TurnOffSpecialize(); break;
      case OP_SPEC2_ON: psi.TurnOnSpecializeV2(); break;
      case OP_SPEC2_OFF: psi.TurnOffSpecializeV2(); break;
      default: fprintf(stderr, "driver: unknown op %d\n", op.kind); exit(2);
    }
  }
  if (psi.IsFusionEnabled()) psi.TurnOffFusion();
}

static bool write_file(const char *fn, const void *p, size_t bytes) {
  FILE *f = fopen(fn, "wb");
  if (!f) return false;
  size_t w = fwrite(p, 1, bytes, f);
  fclose(f);
  return w == bytes;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void trace(const char *what) {  // IQS_DRIVER_TRACE=1: phase timestamps on stderr
  static const bool on = getenv("IQS_DRIVER_TRACE") != nullptr;
  static double t0 = now_s();
  if (on) fprintf(stderr, "[driver %6.3f s] %s\n", now_s() - t0, what);
}

int main(int argc, char **argv) {
  trace("start");
  iqs::mpi::Environment env(argc, argv, false);
  trace("environment ready");
  if (env.IsUsefulRank() == false) return 0;
  if (argc < 2) {
    fprintf(stderr, "usage: %s program.bin [--state-in f] [--state-out f] [--scalars-out f] [--map-out f] [--repeat R]\n", argv[0]);
    return 1;
  }
  const char *state_in = nullptr, *state_out = nullptr, *scalars_out = nullptr, *map_out = nullptr;
  int repeat = 1;
  bool step_norm = true;
  std::vector<std::size_t> step_sizes;
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--state-in" && i + 1 < argc) state_in = argv[++i];
    else if (a == "--state-out" && i + 1 < argc) state_out = argv[++i];
    else if (a == "--scalars-out" && i + 1 < argc) scalars_out = argv[++i];
    else if (a == "--map-out" && i + 1 < argc) map_out = argv[++i];
    else if (a == "--repeat" && i + 1 < argc) repeat = atoi(argv[++i]);
    else if (a == "--no-step-norm") step_norm = false;  // synchronous engines (the reference) need no closing reduction
    else if (a == "--step-sizes" && i + 1 < argc) {  // comma separated op counts, one per timed step
      std::string list = argv[++i];
      for (std::size_t p = 0; p < list.size();) {
        std::size_t q = list.find(',', p);
        if (q == std::string::npos) q = list.size();
        step_sizes.push_back((std::size_t)atol(list.substr(p, q - p).c_str()));
        p = q + 1;
      }
    }
  }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror("program"); return 1; }
  iqs_program_header h;
  if (fread(&h, sizeof(h), 1, f) != 1 || h.magic != IQS_PROGRAM_MAGIC) { fprintf(stderr, "bad program header\n"); return 1; }
  std::vector<iqs_op> ops(h.nops);
  if (h.nops && fread(ops.data(), sizeof(iqs_op), h.nops, f) != h.nops) { fprintf(stderr, "short program\n"); return 1; }
  fclose(f);

  const std::size_t n = h.num_qubits;
  std::size_t tmp = 0;
  if (n > 30) tmp = std::size_t(1) << 30;  // as benchmarks/basic_code_for_scaling.cpp:108-111
  Reg psi(n, h.init == 2 ? "++++" : "base", h.init == 1 ? (std::size_t)h.base_index : 0, tmp);
  if (h.init == 0) {
    if (!state_in) { fprintf(stderr, "init=0 needs --state-in\n"); return 1; }
    FILE *s = fopen(state_in, "rb");
    if (!s) { perror("state-in"); return 1; }
    std::vector<ComplexDP> buf(1 << 16);
    std::size_t L = psi.LocalSize(), first = (std::size_t)iqs::mpi::Environment::GetStateRank() * L;
    fseek(s, (long)(first * sizeof(ComplexDP)), SEEK_SET);
    for (std::size_t i = 0; i < L;) {
      std::size_t want = std::min(buf.size(), L - i);
      if (fread(buf.data(), sizeof(ComplexDP), want, s) != want) { fprintf(stderr, "short state file\n"); return 1; }
      for (std::size_t k = 0; k < want; ++k) psi[i + k] = buf[k];
      i += want;
    }
    fclose(s);
  }

  trace("register initialised");
  std::vector<double> scalars;
  psi.ComputeNorm();  // settle the state in its home memory before the clock starts
  trace("first reduction done");
  auto t0 = std::chrono::steady_clock::now();
  if (!step_sizes.empty()) {
    // timed step by step: one "STEP i seconds" line each (ComputeNorm closes a step so that
    // asynchronous engines have finished it)
    for (std::size_t first = 0, s = 0; first < ops.size() && s < step_sizes.size(); first += step_sizes[s], ++s) {
      std::size_t last = std::min(ops.size(), first + step_sizes[s]);
      std::vector<iqs_op> part(ops.begin() + first, ops.begin() + last);
      auto s0 = std::chrono::steady_clock::now();
      run_ops(psi, part, scalars);
      if (step_norm) psi.ComputeNorm();
      auto s1 = std::chrono::steady_clock::now();
      if (iqs::mpi::Environment::GetStateRank() == 0)
        printf("STEP %zu %.6f\n", s, std::chrono::duration<double>(s1 - s0).count());
    }
  } else {
    for (int r = 0; r < repeat; ++r) {
      if (r) scalars.clear();
      run_ops(psi, ops, scalars);
    }
  }
  double nrm = psi.ComputeNorm();  // forces completion of asynchronous engines
  auto t1 = std::chrono::steady_clock::now();
  double secs = std::chrono::duration<double>(t1 - t0).count();
  // the explicit synchronisation point before ranks read their shards one after the other; engines that
  // keep the state in an internal layout between gates put it back in the documented order here
  iqs::mpi::StateBarrier();
  double secs_settled = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (iqs::mpi::Environment::GetStateRank() == 0)
    printf("TIME %.6f OPS %zu NORM %.15f SETTLED %.6f\n", secs, ops.size() * (size_t)repeat, nrm, secs_settled);

  trace("program done");
  int rank = iqs::mpi::Environment::GetStateRank(), nranks = iqs::mpi::Environment::GetStateSize();
  if (state_out) {
    // every rank writes its shard at its offset (rank 0 first creates the file)
    std::size_t L = psi.LocalSize();
    for (int r = 0; r < nranks; ++r) {
      if (r == rank) {
        FILE *o = fopen(state_out, r == 0 ? "wb" : "r+b");
        if (!o) { perror("state-out"); return 1; }
        fseek(o, (long)((std::size_t)r * L * sizeof(ComplexDP)), SEEK_SET);
        std::vector<ComplexDP> buf(1 << 16);
        for (std::size_t i = 0; i < L;) {
          std::size_t want = std::min(buf.size(), L - i);
          for (std::size_t k = 0; k < want; ++k) buf[k] = psi[i + k];
          fwrite(buf.data(), sizeof(ComplexDP), want, o);
          i += want;
        }
        fclose(o);
      }
      iqs::mpi::StateBarrier();
    }
  }
  trace("state written");
  if (rank == 0 && scalars_out) write_file(scalars_out, scalars.data(), scalars.size() * sizeof(double));
  if (rank == 0 && map_out) {
    std::vector<uint64_t> map(n);
    for (std::size_t i = 0; i < n; ++i) map[i] = psi.qubit_permutation->map[i];
    write_file(map_out, map.data(), map.size() * sizeof(uint64_t));
  }
  return 0;
}

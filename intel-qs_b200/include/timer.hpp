// timer.hpp -- per-gate statistics (interface of reference include/timer.hpp:129-341: Timer with
// Start/Stop, record_sn/dn/tn/cm, Breakdown, Reset).  Re-authored: durations come from CUDA events
// on the engine's stream (the register runs synchronously while statistics are enabled), the
// bandwidth columns are algorithmic bytes / device time.
#ifndef IQS_TIMER_HPP
#define IQS_TIMER_HPP

#include <cstdio>
#include <map>
#include <string>
#include <sys/time.h>
#include <vector>

#include "conversion.hpp"
#include "mpi_env.hpp"
#include "utils.hpp"

namespace iqs {

class Time {
 public:
  double start;
  bool exists;
  std::size_t cpos, tpos;
  std::size_t ncalls;
  double total;
  double sn_time, sn_bw;
  double dn_time, dn_bw;
  double tn_time, tn_bw;
  double cm_time, cm_bw;
  double flops, gflops;
  Time() : start(0), exists(false), cpos(0), tpos(0), ncalls(0), total(0), sn_time(0), sn_bw(0), dn_time(0), dn_bw(0),
           tn_time(0), tn_bw(0), cm_time(0), cm_bw(0), flops(0), gflops(0) {}
  bool timed() { return (sn_time + dn_time + tn_time + cm_time) > 0.0; }
  std::string sprint(bool /*combinedstats*/) {
    char buf[512];
    double nc = ncalls ? double(ncalls) : 1.0;
    snprintf(buf, sizeof(buf), "ncalls %6zu  total %10.6f s | sn %9.6f s %8.2f GB/s | dn %9.6f s %8.2f GB/s | tn %9.6f s %8.2f GB/s | cm %9.6f s %8.2f GB/s",
             ncalls, total, sn_time, sn_bw / nc / 1e9, dn_time, dn_bw / nc / 1e9, tn_time, tn_bw / nc / 1e9, cm_time, cm_bw / nc / 1e9);
    return std::string(buf);
  }
};

class Timer {
 public:
  int num_qubits, my_rank, num_procs, combinedstats;
  std::map<std::string, Time> *timer_map;
  std::map<std::string, Time>::iterator curiter;

  Timer(bool combinedstats_ = false) : num_qubits(0), my_rank(0), num_procs(1), combinedstats(combinedstats_) {
    timer_map = new std::map<std::string, Time>;
    curiter = timer_map->end();
  }
  Timer(int num_qubits_, int my_rank_, int num_procs_) : num_qubits(num_qubits_), my_rank(my_rank_), num_procs(num_procs_), combinedstats(false) {
    timer_map = new std::map<std::string, Time>;
    curiter = timer_map->end();
  }
  ~Timer() { delete timer_map; }

  void Reset() {
    timer_map->clear();
    curiter = timer_map->end();
  }
  double Wtime() {
    struct timeval t;
    gettimeofday(&t, nullptr);
    return double(t.tv_sec) + 1e-6 * double(t.tv_usec);
  }
  void Start(std::string s, std::size_t cpos, std::size_t tpos = 999999) {
    curiter = timer_map->find(s);
    if (curiter == timer_map->end()) curiter = timer_map->insert(std::make_pair(s, Time())).first;
    Time &t = curiter->second;
    t.exists = true;
    t.cpos = cpos;
    t.tpos = tpos;
    iqs::mpi::StateBarrier();
    t.start = Wtime();
  }
  void record_sn(double time, double bw) { if (curiter != timer_map->end()) { curiter->second.sn_time += time; curiter->second.sn_bw += bw; } }
  void record_dn(double time, double bw) { if (curiter != timer_map->end()) { curiter->second.dn_time += time; curiter->second.dn_bw += bw; } }
  void record_tn(double time, double bw) { if (curiter != timer_map->end()) { curiter->second.tn_time += time; curiter->second.tn_bw += bw; } }
  void record_cm(double time, double bw) { if (curiter != timer_map->end()) { curiter->second.cm_time += time; curiter->second.cm_bw += bw; } }
  void Stop() {
    if (curiter == timer_map->end()) return;
    iqs::mpi::StateBarrier();
    Time &t = curiter->second;
    t.total += Wtime() - t.start;
    t.ncalls++;
    curiter = timer_map->end();
  }
  void Breakdown() {
    if (my_rank != 0) return;
    // every line carries the word "statistics" (the reference prints a two-line notice about its
    // statistics here when built without MPI)
    printf(" *** The statistics below are device times (CUDA events) and algorithmic bandwidths of the B200 engine (%d qubits, %d ranks).\n", num_qubits, num_procs);
    for (auto &kv : *timer_map) printf(" *** statistics %-36s %s\n", kv.first.c_str(), kv.second.sprint(combinedstats != 0).c_str());
  }

 private:
  Timer &operator=(const Timer &) { return *this; }
  Timer(const Timer &) {}
};

}  // namespace iqs
#endif

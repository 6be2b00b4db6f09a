// permutation.hpp -- bookkeeping of the qubit order (host only).
// Interface of reference include/permutation.hpp:51-402 (public map/imap, operator[], Find,
// SetNewPermutationFromMap, ExchangeTwoElements, data2program_/program2data_, string helpers,
// ObtainIntemediateInverseMaps); re-authored.  map: element (program qubit) -> position (data
// qubit); imap: position -> element.
#ifndef PERMUTATION_HPP
#define PERMUTATION_HPP

#include <cassert>
#include <cstdio>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "conversion.hpp"
#include "utils.hpp"

namespace iqs {

class Permutation {
 public:
  std::vector<std::size_t> map, imap;
  std::size_t num_elements;

  unsigned operator[](std::size_t i) const { assert(i < num_elements); return (unsigned)map[i]; }
  unsigned operator[](unsigned i) const { assert(i < num_elements); return (unsigned)map[i]; }
  int operator[](int i) const { assert((std::size_t)i < num_elements); return (int)map[i]; }

  std::size_t size() { return map.size(); }

  std::string GetMapStr() { return Join(map); }
  std::string GetImapStr() { return Join(imap); }

  // identity permutation
  Permutation(std::size_t n) : num_elements(n) {
    std::vector<std::size_t> id(n);
    std::iota(id.begin(), id.end(), 0);
    SetNewPermutationFromMap(id, "direct");
  }
  Permutation(std::vector<std::size_t> m, std::string style_of_map = "direct") : num_elements(m.size()) {
    SetNewPermutationFromMap(m, style_of_map);
  }

  // element sitting at `position`
  std::size_t Find(std::size_t position) {
    for (std::size_t e = 0; e < map.size(); ++e)
      if (map[e] == position) return e;
    assert(false && "Permutation::Find: no such position");
    return map.size();
  }

  void SetNewPermutationFromMap(std::vector<std::size_t> m, std::string style_of_map = "direct") {
    assert(m.size() == num_elements);
    std::vector<bool> seen(m.size(), false);
    for (std::size_t v : m) {
      assert(v < m.size());
      seen[v] = true;
    }
    for (bool s : seen) {
      assert(s && "not a permutation");
      (void)s;
    }
    if (style_of_map == "direct") {
      map = m;
      imap.assign(num_elements, 0);
      for (std::size_t e = 0; e < num_elements; ++e) imap[map[e]] = e;
    } else if (style_of_map == "inverse") {
      imap = m;
      map.assign(num_elements, 0);
      for (std::size_t p = 0; p < num_elements; ++p) map[imap[p]] = p;
    } else {
      assert(false && "style_of_map must be 'direct' or 'inverse'");
    }
  }

  void ExchangeTwoElements(std::size_t element_1, std::size_t element_2) {
    std::size_t p1 = map[element_1], p2 = map[element_2];
    map[element_1] = p2;
    map[element_2] = p1;
    imap[p1] = element_2;
    imap[p2] = element_1;
  }

  // "i0 i1 i2 ..." little-endian bit strings
  std::string dec2bin(std::size_t in, std::size_t num_bits) {
    std::string s(num_bits, '0');
    for (std::size_t i = 0; i < num_bits; ++i, in >>= 1)
      if (in & 1) s[i] = '1';
    return s;
  }
  std::size_t bin2dec(std::string in) {
    std::size_t v = 0;
    for (std::size_t i = 0; i < in.size(); ++i)
      if (in[i] == '1') v |= std::size_t(1) << i;
    return v;
  }

  // index in the data representation -> index in the program representation
  inline std::size_t data2program_(std::size_t v) {
    std::size_t r = 0;
    for (std::size_t q = 0; q < num_elements; ++q) r |= ((v >> map[q]) & std::size_t(1)) << q;
    return r;
  }
  // index in the program representation -> index in the data representation
  inline std::size_t program2data_(std::size_t v) {
    std::size_t r = 0;
    for (std::size_t p = 0; p < num_elements; ++p) r |= ((v >> imap[p]) & std::size_t(1)) << p;
    return r;
  }
  std::string data2program(std::size_t v) { return data2program(dec2bin(v, num_elements)); }
  std::string data2program(std::string s) {
    std::string out(s);
    for (std::size_t q = 0; q < s.size(); ++q) out[q] = s[map[q]];
    return out;
  }
  std::string program2data(std::size_t v) { return program2data(dec2bin(v, num_elements)); }
  std::string program2data(std::string s) {
    std::string out(s);
    for (std::size_t p = 0; p < s.size(); ++p) out[p] = s[imap[p]];
    return out;
  }

  void Print() { printf("qubit permutation: %s\n", GetMapStr().c_str()); }
  void PrintRange() {
    for (std::size_t i = 0; i < (UL(1) << num_elements); ++i) printf("map(%3lu) = %3lu\n", i, bin2dec(data2program(i)));
  }

  // Split "current -> target" into (1) a reshuffle of the positions < M, (2) a reshuffle of the
  // positions >= M, (3) pairwise exchanges between a position < M and one >= M.
  // int_1_imap / int_2_imap are the inverse maps after steps (1) and (2).
  void ObtainIntemediateInverseMaps(std::vector<std::size_t> target_map, std::size_t M,
                                    std::vector<std::size_t> &int_1_imap, std::vector<std::size_t> &int_2_imap) {
    assert(M <= num_elements);
    assert(target_map.size() == num_elements);
    // An element whose target lies on the other side of M is parked, within its own side, at the
    // position reached by following the chain target(current occupant) until it returns to this side.
    int_1_imap = imap;
    for (std::size_t pos = 0; pos < M; ++pos) {
      std::size_t p = target_map[imap[pos]];
      while (p >= M) p = target_map[imap[p]];
      int_1_imap[p] = imap[pos];
    }
    int_2_imap = int_1_imap;
    for (std::size_t pos = M; pos < num_elements; ++pos) {
      std::size_t p = target_map[int_1_imap[pos]];
      while (p < M) p = target_map[int_1_imap[p]];
      int_2_imap[p] = int_1_imap[pos];
    }
  }

 private:
  static std::string Join(std::vector<std::size_t> const &v) {
    std::string s;
    for (std::size_t x : v) s += " " + iqs::toString(x);
    return s;
  }
};

}  // namespace iqs
#endif

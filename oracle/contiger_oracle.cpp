// oracle/contiger_oracle.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Sequential, deterministic restatement of the reference's Contiger (src/contig_assembly.cpp,
// live parts only) over the oracle CQF model: seeds from read middles (:1839-1924 worker rule,
// :2049-2149 master rule), get_unitig_forward (:3028-3218), insert_or_replace (:3018-3025),
// check_unitig (:935-954), track_kmer_worker (:956-1010), build_graph_worker (:1012-1084) and the
// FASTA writer (:603-629).  TBB containers become std:: containers, the job queue is drained after
// every seed (FIFO or LIFO, to show the unitig set does not depend on the visiting order).
//
// Parity pinning: the reference Contiger cannot be built here (boost, TBB); its output is not even
// run-to-run deterministic at "island" level.  This restatement is checked
// against hand-built graphs (tests/test_contiger_oracle.py) and against its own invariance under
// visiting order -- "parity unpinned" against reference output, and it says so in DESIGN.md.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <string>
#include <unordered_map>
#include <vector>

#include "shq_oracle.h"

namespace {

const char DNA_bases[4] = {'A', 'C', 'G', 'T'};   // base/global.h:110

char rc_base(char b) {   // base/Utility.h:96-118
  switch (b) { case 'A': case 'a': return 'T'; case 'T': case 't': return 'A'; case 'G': case 'g': return 'C'; case 'C': case 'c': return 'G'; }
  return 'T';  // unreachable for 2-bit strings
}
std::string rc(const std::string& s) {
  std::string r(s.rbegin(), s.rend());
  for (auto& c : r) c = rc_base(c);
  return r;
}
// DNAString stores 2 bits per base: anything that is not ACGT/acgt becomes 'A' (base/DNA_string.h:18-40)
std::string to_2bit(const std::string& s) {
  std::string r = s;
  for (auto& c : r) {
    switch (c) { case 'A': case 'a': c = 'A'; break; case 'C': case 'c': c = 'C'; break; case 'G': case 'g': c = 'G'; break;
                 case 'T': case 't': c = 'T'; break; default: c = 'A'; }
  }
  return r;
}

struct Contig { std::string seq; int median_abundance; };

double median(std::vector<int>& nums) {   // base/Utility.cpp:27-40
  if (nums.empty()) return 0;
  if (nums.size() == 1) return nums[0];
  std::sort(nums.begin(), nums.end());
  size_t t = nums.size() / 2;
  return nums.size() % 2 == 0 ? (nums[t - 1] + nums[t]) / 2.0 : nums[t];
}

struct Ctx {
  void* cqf; int K, s_min, x_min, x_max; int hb;
  std::vector<Contig> contigs;
  std::unordered_map<std::string, int> start2unitig;
  std::deque<int> jobs;
  uint64_t n_queries = 0;

  uint64_t key_of(const std::string& kmer) {
    uint64_t fh, rh;
    orc_ntpc64_base(kmer.c_str(), (unsigned)K, &fh, &rh);
    uint64_t h = fh < rh ? fh : rh;
    return hb >= 64 ? h : (h & ((1ULL << hb) - 1));
  }
  bool count_set_traveled(const std::string& kmer, uint64_t* count) {
    n_queries++;
    return orc_cqf_query_set_traveled(cqf, key_of(kmer), count) != 0;
  }
  bool insert_or_replace(const std::string& k, int idx) {   // :3018-3025
    auto it = start2unitig.find(k);
    if (it == start2unitig.end()) { start2unitig[k] = idx; return true; }
    if (it->second >= idx) { it->second = idx; return true; }
    return false;
  }

  void get_unitig_forward(int cid) {   // :3028-3218
    std::string& contig_seq = contigs[cid].seq;
    std::string first_kmer = contig_seq.substr(0, K);
    std::string current_kmer = contig_seq.substr(contig_seq.length() - K);
    std::string current_kmer_RC = rc(current_kmer);
    std::vector<int> abundances(contig_seq.length() - K + 1, contigs[cid].median_abundance);
    for (;;) {
      std::string fix = current_kmer.substr(1);
      bool cand_before[4] = {false, false, false, false}, cand_after[4] = {false, false, false, false};
      uint64_t ab_before[4] = {0, 0, 0, 0}, ab_after[4] = {0, 0, 0, 0};
      int cb = 0, ca = 0, nb = 0, na = 0;
      for (int x = 0; x < 4; x++) {                              // successors fix+x  (:3064-3092)
        std::string kmer = fix + DNA_bases[x];
        uint64_t cnt;
        bool traveled = count_set_traveled(kmer, &cnt);
        if (cnt >= (uint64_t)s_min) {
          if (traveled && start2unitig.count(kmer)) { na++; ab_after[x] = cnt; }
          else { ab_after[x] = cnt; cand_after[x] = true; ca++; }
        }
      }
      std::string kmer = current_kmer_RC;                        // predecessors, queried as RC(fix)+x (:3094-3120)
      for (int x = 0; x < 4; x++) {
        if (DNA_bases[x] == current_kmer_RC[K - 1]) continue;
        kmer[K - 1] = DNA_bases[x];
        uint64_t cnt;
        bool traveled = count_set_traveled(kmer, &cnt);
        if (cnt >= (uint64_t)s_min) {
          if (traveled && start2unitig.count(kmer)) { nb++; }
          else { ab_before[x] = cnt; cand_before[x] = true; cb++; }
        }
      }
      if ((nb + cb) || (na + ca) > 1) {                           // branch (:3122-3166)
        if (!insert_or_replace(current_kmer_RC, cid)) { contigs[cid].seq.clear(); break; }
        contigs[cid].median_abundance = (int)median(abundances);
        for (int x = 0; x < 4; x++) {
          if (!cand_after[x]) continue;
          std::string k2 = fix + DNA_bases[x];
          if (!start2unitig.count(k2)) {
            contigs.push_back(Contig{k2, (int)ab_after[x]});
            start2unitig[k2] = (int)contigs.size() - 1;
            jobs.push_back((int)contigs.size() - 1);
          }
        }
        std::string k3 = current_kmer_RC;
        for (int x = 0; x < 4; x++) {
          if (!cand_before[x]) continue;
          k3[K - 1] = DNA_bases[x];
          if (!start2unitig.count(k3)) {
            contigs.push_back(Contig{k3, (int)ab_before[x]});
            start2unitig[k3] = (int)contigs.size() - 1;
            jobs.push_back((int)contigs.size() - 1);
          }
        }
        break;
      } else if (ca == 1) {                                      // linear extension (:3167-3191)
        int x = 0;
        for (x = 0; x < 4; x++) if (cand_after[x]) break;
        current_kmer = fix + DNA_bases[x];
        if (current_kmer == first_kmer) {                        // pure circle
          if (!insert_or_replace(first_kmer, cid) || !insert_or_replace(current_kmer_RC, cid)) contigs[cid].seq.clear();
          else contigs[cid].median_abundance = (int)median(abundances);
          break;
        }
        current_kmer_RC = rc_base(DNA_bases[x]) + current_kmer_RC.substr(0, K - 1);
        contigs[cid].seq += DNA_bases[x];
        abundances.push_back((int)ab_after[x]);
        continue;
      } else {                                                   // sole successor is a node, or dead end (:3192-3214)
        if (!insert_or_replace(current_kmer_RC, cid)) contigs[cid].seq.clear();
        else contigs[cid].median_abundance = (int)median(abundances);
        break;
      }
    }
  }

  void drain(bool lifo) {
    while (!jobs.empty()) {
      int j;
      if (lifo) { j = jobs.back(); jobs.pop_back(); } else { j = jobs.front(); jobs.pop_front(); }
      get_unitig_forward(j);
    }
  }

  // processDataChunk (:1839-1924) / master loop (:2049-2149): the line cursor is restated literally,
  // including the `continue`s that skip the two skipLines().
  void process_chunk(const char* text, size_t n, bool master_rule, bool lifo) {
    size_t p = 0;
    auto read_line = [&](std::string& out) -> bool {
      if (p == n) return false;
      const char* nl = (const char*)memchr(text + p, '\n', n - p);
      size_t e = nl ? (size_t)(nl - text) : n;
      out.assign(text + p, e - p);
      p = nl ? e + 1 : n;
      return true;
    };
    auto skip_lines = [&](int k) { std::string t; while (k--) if (!read_line(t)) return; };
    std::string line, seq;
    while (read_line(line)) {
      if (line.empty()) continue;
      if (line[0] != '@') continue;
      if (!read_line(seq)) break;
      if ((int)seq.length() < K) { skip_lines(2); continue; }
      int seq_len = (int)seq.length();
      int middle = master_rule ? seq_len / 2 : seq_len / 2 - K / 2;
      if (seq_len >= K && middle <= seq_len - K) {
        std::string kmer = seq.substr(middle, K);
        for (auto& c : kmer) if (c > 0x60) c -= 32;              // to_upper_DNA, base/Utility.h:120-155
        if (kmer.find_first_of("nN") != std::string::npos) continue;
        uint64_t cnt;
        if (count_set_traveled(kmer, &cnt)) continue;
        if (cnt < (uint64_t)x_min || cnt > (uint64_t)x_max) continue;
        std::string k2 = to_2bit(kmer);
        contigs.push_back(Contig{k2, (int)cnt});
        int cid = (int)contigs.size() - 1;
        get_unitig_forward(cid);
        if (!contigs[cid].seq.empty()) {
          auto it = start2unitig.find(k2);
          if (it != start2unitig.end()) {
            if (it->second > cid) { contigs[cid].seq = rc(contigs[cid].seq); get_unitig_forward(cid); }
            else if (it->second < cid) contigs[cid].seq.clear();
          } else {
            contigs[cid].seq = rc(contigs[cid].seq);
            get_unitig_forward(cid);
          }
        }
        drain(lifo);
      }
      skip_lines(2);
    }
  }
};

struct Result {
  std::vector<std::string> seqs; std::vector<int> km;
  std::vector<std::vector<int>> after, before;
  std::string fasta;
  uint64_t n_queries = 0;
};

}  // namespace

extern "C" {

// Runs the whole Contiger over in-memory read files against the oracle CQF `cqf` (its traveled
// bits are modified).  flags: bit0 = master seed rule (len/2) instead of the worker rule
// (len/2 - K/2); bit1 = LIFO job order.  Returns a result handle.
void* orc_contiger_run(void* cqf, int hb, int nfiles, const char* const* bufs, const size_t* lens, uint64_t part_size,
                       int K, int s_min, int x_min, int x_max, int flags) {
  Ctx c;
  c.cqf = cqf; c.K = K; c.s_min = s_min; c.x_min = x_min; c.x_max = x_max; c.hb = hb;
  c.contigs.resize(1);
  void* ch = orc_chunker_new(nfiles, bufs, lens, part_size);
  std::vector<char> buf((part_size ? part_size : (1ULL << 23)) + 65536);
  for (;;) {
    int64_t n = orc_chunker_next(ch, buf.data(), buf.size(), nullptr);
    if (n < 0) break;
    c.process_chunk(buf.data(), (size_t)n, flags & 1, flags & 2);
  }
  orc_chunker_free(ch);
  // check_unitig (:935-954)
  for (size_t id = 1; id < c.contigs.size(); id++) {
    if (c.contigs[id].seq.empty()) continue;
    auto it = c.start2unitig.find(c.contigs[id].seq.substr(0, K));
    if (it != c.start2unitig.end() && it->second != (int)id) c.contigs[id].seq.clear();
  }
  // track_kmer_worker (:956-1010)
  Result* r = new Result();
  int counter = 1;
  for (size_t id = 1; id < c.contigs.size(); id++) {
    const std::string& s = c.contigs[id].seq;
    if (s.empty()) continue;
    std::string first = s.substr(0, K), last_rc = rc(s.substr(s.length() - K));
    if (first == last_rc) { auto it = c.start2unitig.find(first); if (it != c.start2unitig.end()) it->second = counter; }
    else {
      auto it = c.start2unitig.find(last_rc); if (it != c.start2unitig.end()) it->second = -counter;
      it = c.start2unitig.find(first); if (it != c.start2unitig.end()) it->second = counter;
    }
    r->seqs.push_back(s); r->km.push_back(c.contigs[id].median_abundance);
    counter++;
  }
  // build_graph_worker (:1012-1084)
  r->after.resize(r->seqs.size()); r->before.resize(r->seqs.size());
  for (size_t u = 0; u < r->seqs.size(); u++) {
    const std::string& s = r->seqs[u];
    std::string fix = s.substr(s.length() - K + 1, K - 1);
    for (int x = 0; x < 4; x++) { auto it = c.start2unitig.find(fix + DNA_bases[x]); if (it != c.start2unitig.end()) r->after[u].push_back(it->second); }
    fix = rc(s.substr(0, K - 1));
    for (int x = 3; x >= 0; x--) { auto it = c.start2unitig.find(fix + DNA_bases[x]); if (it != c.start2unitig.end()) r->before[u].push_back(it->second); }
  }
  // FASTA (:603-629)
  for (size_t u = 0; u < r->seqs.size(); u++) {
    const std::string& s = r->seqs[u];
    r->fasta += ">" + std::to_string(u) + " LN:i:" + std::to_string(s.length()) + " KC:i:" +
                std::to_string((long long)r->km[u] * (long long)(s.length() - K + 1)) + " km:f:" + std::to_string(r->km[u]);
    for (int t : r->after[u]) r->fasta += t > 0 ? " L:+:" + std::to_string(t - 1) + ":+" : " L:+:" + std::to_string(-t - 1) + ":-";
    for (int t : r->before[u]) r->fasta += t > 0 ? " L:-:" + std::to_string(t - 1) + ":+" : " L:-:" + std::to_string(-t - 1) + ":-";
    r->fasta += "\n" + s + "\n";
  }
  r->n_queries = c.n_queries;
  return r;
}
uint64_t orc_contiger_fasta_size(void* r) { return ((Result*)r)->fasta.size(); }
void orc_contiger_fasta(void* r, char* dst) { memcpy(dst, ((Result*)r)->fasta.data(), ((Result*)r)->fasta.size()); }
uint64_t orc_contiger_queries(void* r) { return ((Result*)r)->n_queries; }
void orc_contiger_free(void* r) { delete (Result*)r; }

}  // extern "C"

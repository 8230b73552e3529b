// Contiger -- drop-in command line of the reference's src/contig_assembly.cpp (main: 191-793, live parts
// 246-303 and 586-630) over libshq.so.  Same options and defaults (:23-75), same progress lines on stdout,
// same output format (BCALM/Minia unitig FASTA with LN/KC/km/L tags, :603-629).  The graph work runs on the
// GPU through the C ABI (shq_unitigs_*).
#include <time.h>

#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/shq.h"
#include "cli_options.h"

using namespace std;

static void DisplayCurrentDateTime() {   // base/Utility.h:695-703
  time_t now = time(0);
  struct tm tstruct = *localtime(&now);
  char buf[80];
  strftime(buf, sizeof(buf), "%Y-%m-%d %X", &tstruct);
  cout << buf << endl;
}

int main(int argc, char* argv[]) {
  Options opt;
  opt.add("help", 'h', false, false, nullptr, "print help messages");
  opt.add("", 'k', true, true, nullptr, "k-mer size");
  opt.add("input", 'i', true, true, nullptr, "a file containing a list of read file name(s), should be absolute address if the read files are not in the running directory");
  opt.add("format", 'f', true, false, "f", "format of the input: g(gzip); b(bzip2); f(plain fastq)");
  opt.add("cqf", 'c', true, true, nullptr, "the counting quotient filter built with the same 'k'");
  opt.add("abundance_min", 's', true, false, "2", "minimum coverage of k-mers used to extend the assembly");
  opt.add("solid_abundance_min", 'x', true, false, "2", "minimum coverage of a solid k-mer to start the assembly");
  opt.add("solid_abundance_max", 'X', true, false, "1000000", "maximum coverage of a solid k-mer to start the assembly");
  opt.add("", 't', true, false, "16", "number of threads");
  opt.add("output", 'o', true, false, "unitigs.fa", "output contig file name (fasta)");
  opt.add("device", 0, true, false, "0", "CUDA device ordinal (extension; the reference has no such option)");
  string desc = opt.describe(string(argv[0]) + "  <options>\nOptions:");
  try {
    opt.parse(argc, argv);
  } catch (const exception& e) {
    cerr << "terminate called after throwing an instance of 'boost::program_options::error'\n  what():  " << e.what() << endl;
    return 134;
  }
  if (argc == 1 || opt.count("help")) { cerr << desc << "\n"; return 0; }
  try {
    opt.notify();
  } catch (const exception& e) {
    cerr << "terminate called after throwing an instance of 'boost::program_options::required_option'\n  what():  " << e.what() << endl;
    return 134;
  }
  shq_unitig_opts uo;
  uo.K = stoi(opt.get("-k"));
  uo.abundance_min = stoi(opt.get("abundance_min"));
  uo.solid_abundance_min = stoi(opt.get("solid_abundance_min"));
  uo.solid_abundance_max = stoi(opt.get("solid_abundance_max"));
  int thread_num = stoi(opt.get("-t"));
  char fmode = opt.get("format").empty() ? 'f' : opt.get("format")[0];
  if (fmode != 'g' && fmode != 'b' && fmode != 'f') {
    std::cerr << "[Error] unrecognized file format " << fmode << endl;
    fmode = 'f';
  }
  cout << "Params:" << endl                                  // base/Params.h:63-71
       << "\tkmer size:                " << uo.K << endl
       << "\tkmer min. abundance:      " << uo.abundance_min << endl
       << "\tsolid kmer min. abundance:" << uo.solid_abundance_min << endl
       << "\tsolid kmer max. abundance:" << uo.solid_abundance_max << endl
       << "\tthreads:                  " << thread_num << endl;

  vector<string> seqFileNames;
  ifstream fin(opt.get("input"));
  string line;
  while (getline(fin, line)) {
    if (line.empty()) continue;
    seqFileNames.push_back(line);
  }
  vector<const char*> paths;
  for (auto& f : seqFileNames) paths.push_back(f.c_str());
  shq_reader_t rd = nullptr;
  int rc = shq_reader_open(paths.data(), (int)paths.size(), fmode, &rd);
  if (rc) { cerr << "cannot open inputs: " << shq_strerror(rc) << endl; return 1; }

  DisplayCurrentDateTime();
  cout << "[CQF] load cqf from disk" << endl;
  shq_t h = nullptr;
  rc = shq_load(opt.get("cqf").c_str(), stoi(opt.get("device")), &h);
  if (rc) { perror("Error opening file for deserializing\n"); return EXIT_FAILURE; }
  cout << "[CQF] cqf loaded!" << endl;

  DisplayCurrentDateTime();
  cout << "[Unitig] find unitigs" << endl << std::flush;
  shq_unitigs_t u = nullptr;
  rc = shq_unitigs_begin(h, &uo, &u);
  if (rc) { cerr << "unitig setup failed: " << shq_strerror(rc) << ": " << shq_last_error(h) << endl; return 1; }
  for (;;) {
    const char* chunk = nullptr; size_t n = 0;
    rc = shq_reader_next(rd, &chunk, &n, nullptr);
    if (rc) { cerr << "Error: Wrong input file!" << endl; return EXIT_FAILURE; }
    if (!chunk) break;
    if (n == 0) continue;
    rc = shq_unitigs_add_reads(u, chunk, n);
    if (rc) { cerr << "reading failed: " << shq_strerror(rc) << ": " << shq_last_error(h) << endl; return 1; }
  }
  shq_reader_close(rd);
  DisplayCurrentDateTime();
  cout << "[Check] make sure that there is no duplicate seq." << endl << std::flush;
  rc = shq_unitigs_build(u);
  if (rc) { cerr << "unitig construction failed: " << shq_strerror(rc) << ": " << shq_last_error(h) << endl; return 1; }
  size_t n_unitigs = 0, total = 0;
  shq_unitigs_count(u, &n_unitigs, &total);
  cout << "[Unitig] " << n_unitigs << " unitigs reported of length " << total << " bp in total" << endl;
  cout << "[Unitig] build unitig graph." << endl;
  DisplayCurrentDateTime();
  cout << "[Dump] save the unitig graph to file." << endl;
  DisplayCurrentDateTime();
  rc = shq_unitigs_write_fasta(u, opt.get("output").c_str());
  if (rc) { cerr << "cannot write " << opt.get("output") << endl; return 1; }
  shq_unitigs_free(u);
  shq_destroy(h);
  return 0;
}

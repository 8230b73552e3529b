// CQF-deNoise -- drop-in command line of the reference's src/CQF-deNoise.cpp:53-224 over libshq.so.
// Same options, defaults, banner and output file; the counting itself runs on the GPU through the
// C ABI (include/shq.h).  Host logic mirrors main(): list-file handling (:59-81), parameter block
// (:83-161 via shq_param_block), banner (:185-210), build + save (:212-221).
#include <time.h>

#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/shq.h"
#include "cli_options.h"

using namespace std;

static string currentDateTime() {   // base/global.cpp:14-21
  time_t now = time(0);
  struct tm tstruct = *localtime(&now);
  char buf[80];
  strftime(buf, sizeof(buf), "%Y-%m-%d.%X", &tstruct);
  return buf;
}

int main(int argc, char* argv[]) {
  Options opt;
  opt.add("help", 'h', false, false, nullptr, "print help messages");
  opt.add("", 'k', true, true, nullptr, "k-mer size");
  opt.add("trueKmer", 'n', true, true, nullptr, "number of unique true k-mers");
  opt.add("", 'N', true, true, nullptr, "total number of k-mers");
  opt.add("alpha", 'e', true, false, "-1", "average base error rate, when specified, the <errorProfile> is ignored");
  opt.add("errorProfile", 0, true, false, "", "error profile in a file, each line with error rate for the corresponding base, e.g. the error rate of the second base is specified in the second line");
  opt.add("fr", 0, true, false, "0", "tolerable rate of true k-mers being wrongly removed, default: 1/<trueKmer>");
  opt.add("deNoise", 0, true, false, "-1", "number of rounds of deNoise, when specified, the <fr> is ignored");
  opt.add("endDeNoise", 0, false, false, nullptr, "call deNoise after processing all the k-mers (not counted into the <deNoise>)");
  opt.add("", 't', true, false, "16", "number of threads");
  opt.add("format", 'f', true, true, nullptr, "format of the input: g(gzip); b(bzip2); f(plain fastq)");
  opt.add("input", 'i', true, true, nullptr, "a file containing list of input file name(s), should be in the same directory as the fastq file(s)");
  opt.add("output", 'o', true, false, nullptr, "output file name");
  opt.add("device", 0, true, false, "0", "CUDA device ordinal (extension; the reference has no such option)");
  string desc = opt.describe(string(argv[0]) + "  <options>\nOptions");
  try {
    opt.parse(argc, argv);
  } catch (const exception& e) {
    cerr << "terminate called after throwing an instance of 'boost::program_options::error'\n  what():  " << e.what() << endl;
    return 134;
  }
  if (argc == 1 || opt.count("help")) { cerr << endl << desc << "\n"; return 0; }
  if (stod(opt.get("alpha")) == -1 && opt.get("errorProfile") == "") {
    cerr << endl << "Please specify either <alpha> or <errorProfile>" << endl << endl;
    cerr << desc << "\n";
    return 0;
  }
  try {
    opt.notify();
  } catch (const exception& e) {
    cerr << "terminate called after throwing an instance of 'boost::program_options::required_option'\n  what():  " << e.what() << endl;
    return 134;
  }

  vector<string> files;
  string flist = opt.get("input");
  string file_prefix = "";
  auto pos = flist.find_last_of("/\\");
  if (pos != string::npos) file_prefix = flist.substr(0, pos + 1);
  ifstream fin(flist);
  string line;
  if (fin.is_open()) {
    while (!fin.eof()) {
      getline(fin, line);
      if (line.empty()) continue;
      files.push_back(file_prefix + line);
    }
    fin.close();
  } else {
    cerr << "Failed to open file: " << flist << endl;
    return 0;
  }

  uint32_t seed = 2038074761;
  uint16_t thread_num = (uint16_t)stoi(opt.get("-t"));
  int K = stoi(opt.get("-k"));
  double alpha = stod(opt.get("alpha"));
  string errorProfile = opt.get("errorProfile");
  uint64_t n_true_kmers = stoull(opt.get("trueKmer"));
  uint64_t total_kmers = stoull(opt.get("-N"));
  int num_deNoise = stoi(opt.get("deNoise"));
  double fr = stod(opt.get("fr"));
  bool end_deNoise = opt.count("endDeNoise");

  shq_params p;
  int rc = shq_param_block(K, n_true_kmers, total_kmers, alpha, errorProfile.c_str(), num_deNoise, fr, &p);
  if (rc) { cerr << "bad parameters: " << shq_strerror(rc) << endl; return 1; }
  uint64_t qb = p.qb, hb = p.hb;
  num_deNoise = p.num_deNoise;
  fr = p.fr;

  string output_file;
  if (!opt.count("output")) output_file = "k" + to_string(K) + ".t" + to_string(thread_num) + ".s" + to_string(qb) + ".ser";
  else output_file = opt.get("output");

  char fmode = opt.get("format").empty() ? '?' : opt.get("format")[0];
  if (fmode != 'g' && fmode != 'b' && fmode != 'f') {
    cerr << "Unrecognized file type " << fmode << endl;
    cerr << "run following to get help" << endl;
    cerr << "\t" << argv[0] << " --help" << endl;
    return 0;
  }

  double mean = (double)(p.num_true_kmers / n_true_kmers);
  cerr << "CQF-deNoise settings:" << endl
       << "qb: " << qb << endl
       << "hb: " << hb << endl
       << "thread_num: " << thread_num << endl
       << "K: " << K << endl
       << "number of true k-mers: " << n_true_kmers << endl
       << "tolerable wrong removal rate: " << fr << endl
       << "number of deNoise rounds: " << num_deNoise << endl
       << "deNoise after processing all k-mers: " << (end_deNoise ? "true" : "false") << endl
       << "number of unique k-mers triggering deNoise: " << p.trigger << endl;
  cerr << "#deNoise rounds leading to the same size of CQF: [" << p.lower_bound << ", "
       << (p.upper_bound == 0 ? string("+oo") : to_string(p.upper_bound)) << "]" << endl;
  cerr << "wrong removal rate leading to the same #deNoise rounds: [" << shq_cdfpoi_positive(p.lower_bound, mean) << ", ";
  if (p.upper_bound == 0) cerr << "1";
  else cerr << shq_cdfpoi_positive(p.upper_bound, mean);
  cerr << "]" << endl << endl;

  shq_t h = nullptr;
  rc = shq_create((int)qb, (int)hb, seed, stoi(opt.get("device")), &h);
  if (rc) { cerr << "shq_create failed: " << shq_strerror(rc) << endl; return 1; }
  shq_set_denoise_schedule(h, p.trigger, (uint32_t)num_deNoise, end_deNoise ? 1 : 0);
  shq_set_progress_log(h, 1, n_true_kmers);

  time_t start_time = time(NULL);
  cerr << currentDateTime() << endl;
  cerr << "Start to build K-mer spectrum..." << endl;
  vector<const char*> paths;
  for (auto& f : files) paths.push_back(f.c_str());
  shq_reader_t rd = nullptr;
  rc = shq_reader_open(paths.data(), (int)paths.size(), fmode, &rd);
  if (rc) { cerr << "cannot open inputs: " << shq_strerror(rc) << endl; return 1; }
  time_t t0 = time(NULL);
  rc = shq_build_kmer_spectrum(h, rd, K);
  shq_reader_close(rd);
  if (rc) { cerr << "counting failed: " << shq_strerror(rc) << ": " << shq_last_error(h) << endl; return 1; }
  cerr << "Time for building K-mer spectrum without dumping to disk: " << difftime(time(NULL), t0) << " seconds." << endl;
  uint64_t ndistinct = 0, nelts = 0;
  shq_counts(h, &ndistinct, &nelts);
  uint32_t n_deNoise = (uint32_t)num_deNoise;
  cerr << "Estimated probability of true k-mers with freq<=" << (end_deNoise ? n_deNoise + 1 : n_deNoise) << " is: "
       << shq_cdfpoi((end_deNoise ? n_deNoise + 1 : n_deNoise), (nelts / (double)ndistinct)) << endl;
  rc = shq_save(h, output_file.c_str());
  if (rc) { cerr << "Error opening file for serializing\n: " << shq_last_error(h) << endl; return 1; }
  cerr << "Finished building K-mer spectrum!" << endl;
  double seconds = difftime(time(NULL), start_time);
  cerr << "Time for building K-mer spectrum: " << seconds << " seconds." << endl;
  shq_destroy(h);
  return 0;
}

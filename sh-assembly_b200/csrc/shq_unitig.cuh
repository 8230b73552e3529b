// shq_unitig.cuh -- Contiger device path (K6-K8).  Placeholder until the counting path is
// parity-green on the GPU: the entry points exist so that the ABI is complete, and report
// SHQ_ERR_STATE instead of silently doing nothing.
#pragma once
extern "C" {
struct shq_unitigs_s { shq_t h; shq_unitig_opts opts; };
int shq_unitigs_begin(shq_t h, const shq_unitig_opts* opts, shq_unitigs_t* out) {
  if (!h || !opts || !out) return SHQ_ERR_ARG;
  *out = nullptr;
  h->last_error = "unitig path not built yet";
  return SHQ_ERR_STATE;
}
int shq_unitigs_add_reads(shq_unitigs_t, const char*, size_t) { return SHQ_ERR_STATE; }
int shq_unitigs_build(shq_unitigs_t) { return SHQ_ERR_STATE; }
int shq_unitigs_count(shq_unitigs_t, size_t*, size_t*) { return SHQ_ERR_STATE; }
int shq_unitigs_write_fasta(shq_unitigs_t, const char*) { return SHQ_ERR_STATE; }
int shq_unitigs_fasta(shq_unitigs_t, char*, size_t, size_t*) { return SHQ_ERR_STATE; }
void shq_unitigs_free(shq_unitigs_t u) { delete u; }
}

// Output writers of the sampling CLI, host C++:
//
//   * per-chain angle tables as pandas-compatible csv.gz  - replaces `DataFrame.to_csv(".../generated_{i}.csv.gz")`,
//     /root/reference/bin/sample.py:365-370 (and the --fullhistory snapshots, :372-385)
//   * backbone PDB files (N, CA, C of GLY residues, chain A)  - replaces angles_and_coords.write_coords_to_pdb,
//     /root/reference/foldingdiff/angles_and_coords.py:187-253 (biotite Atom fields: res_name GLY, occupancy 1.0,
//     b_factor 5.0, element N / C / C, atom_id = running index from 1, res_id from 1)
//
// The reference fans the per-chain Python work out over a multiprocessing pool (bin/sample.py:105-128); here a
// batch call formats and compresses chains on a small std::thread pool and touches the Python interpreter once.
//
// Number formatting is numpy's str(float32) - what pandas writes for a float32 frame: shortest digits that
// round-trip, positional notation with at least one fractional digit for 1e-4 <= |x| < 1e6 (and for 0), scientific
// "d.ddde-XX" otherwise; nan / inf spelled like pandas ("", "inf", "-inf").
#pragma once
#include <zlib.h>

#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace fdw {

inline void append_f32(std::string& out, float v) {
  if (std::isnan(v)) return;  // pandas writes NaN as the empty field
  if (std::isinf(v)) { out += v < 0 ? "-inf" : "inf"; return; }
  char buf[64];
  const double a = std::fabs((double)v);  // numpy compares the widened value: float32(1e-4) < 1e-4 prints as "1e-04"
  if (a == 0.0 || (a >= 1e-4 && a < 1e6)) {
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    out.append(buf, r.ptr);
    if (!memchr(buf, '.', (size_t)(r.ptr - buf))) out += ".0";
  } else {
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    out.append(buf, r.ptr);
  }
}

// index column + header like DataFrame.to_csv(): ",phi,psi,...\n0,v,v,...\n"
inline std::string angles_csv_text(const float* angles, int n_rows, int n_features, int row_stride,
                                   const char* const* names) {
  std::string s;
  s.reserve((size_t)n_rows * (size_t)n_features * 12 + 128);
  for (int f = 0; f < n_features; ++f) { s += ','; s += names[f]; }
  s += '\n';
  for (int r = 0; r < n_rows; ++r) {
    s += std::to_string(r);
    for (int f = 0; f < n_features; ++f) { s += ','; append_f32(s, angles[(size_t)r * row_stride + f]); }
    s += '\n';
  }
  return s;
}

inline int write_gz(const char* path, const std::string& text, int level) {
  gzFile f = gzopen(path, level >= 0 && level <= 9 ? (std::string("wb") + char('0' + level)).c_str() : "wb");
  if (!f) return 1;
  size_t off = 0;
  while (off < text.size()) {
    const unsigned chunk = (unsigned)std::min<size_t>(text.size() - off, 1u << 30);
    const int w = gzwrite(f, text.data() + off, chunk);
    if (w <= 0) { gzclose(f); return 2; }
    off += (size_t)w;
  }
  return gzclose(f) == Z_OK ? 0 : 3;
}

// coords: [3 * n_res][3] (N, CA, C per residue), Angstrom.  PDB format v3.3 ATOM records (80 columns) + CONECT.
inline std::string backbone_pdb_text(const float* coords, int n_atoms) {
  static const char* kName[3] = {" N  ", " CA ", " C  "};
  static const char* kElem[3] = {" N", " C", " C"};
  std::string s;
  s.reserve((size_t)n_atoms * 81 + (size_t)n_atoms / 3 * 34 + 16);
  char line[96];
  for (int i = 0; i < n_atoms; ++i) {
    const float* p = coords + (size_t)i * 3;
    const int n = snprintf(line, sizeof(line), "ATOM  %5d %s %3s %c%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s  \n", i + 1,
                           kName[i % 3], "GLY", 'A', i / 3 + 1, (double)p[0], (double)p[1], (double)p[2], 1.0, 5.0, kElem[i % 3]);
    s.append(line, (size_t)n);
  }
  // the reference adds a single bond between consecutive atoms (angles_and_coords.py:236-240); biotite's PDB writer
  // spells out only the inter-residue ones (C of residue i - N of residue i + 1), both directions, as CONECT records
  // (golden file: plots/pdb_structures/noising_visualization/fully_noised.pdb of the reference)
  for (int c = 3; c < n_atoms; c += 3) {
    int n = snprintf(line, sizeof(line), "CONECT%5d%5d\n", c, c + 1);
    s.append(line, (size_t)n);
    n = snprintf(line, sizeof(line), "CONECT%5d%5d\n", c + 1, c);
    s.append(line, (size_t)n);
  }
  return s;
}

inline int write_text(const char* path, const std::string& text) {
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  const size_t w = fwrite(text.data(), 1, text.size(), f);
  return (fclose(f) == 0 && w == text.size()) ? 0 : 2;
}

struct BatchJob {
  const float* angles; int n_rows, n_features, row_stride; const char* const* names; const char* csv_path;  // csv_path may be null
  const float* coords; int n_atoms; const char* pdb_path;                                                    // pdb_path may be null
};

// returns the number of failed files
inline int run_batch(const std::vector<BatchJob>& jobs, int n_threads, int gz_level) {
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  auto work = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= jobs.size()) return;
      const BatchJob& j = jobs[i];
      if (j.csv_path && write_gz(j.csv_path, angles_csv_text(j.angles, j.n_rows, j.n_features, j.row_stride, j.names), gz_level)) failed++;
      if (j.pdb_path && write_text(j.pdb_path, backbone_pdb_text(j.coords, j.n_atoms))) failed++;
    }
  };
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > jobs.size()) n_threads = (int)jobs.size();
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return failed.load();
}

}  // namespace fdw

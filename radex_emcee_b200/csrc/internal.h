// Internal declarations shared by the host-side LAMDA loader and the CUDA translation unit.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "radex_b200.h"

#define RB_MAXPART 7
#define RB_MAXLEV 64   // kernels keep one model's rate matrix on chip; LAMDA CO has 41 levels

// Host-side molecular table, SoA.  Level/line indices are 0-based here.
struct rb_mol {
  std::string name;
  double amass = 0.0;
  int nlev = 0, nline = 0, npart = 0;
  std::vector<double> eterm, gstat;                  // [nlev]
  std::vector<int> iupp, ilow;                       // [nline]
  std::vector<double> aeinst, spfreq, eup, xnu;      // [nline]
  struct Partner {
    int id = 0;                                      // LAMDA id 1..7
    int ncoll = 0, ntemp = 0;
    std::vector<double> temps;                       // [ntemp]
    std::vector<int> lcu, lcl;                       // [ncoll]
    std::vector<double> rates_tc;                    // [ntemp][ncoll]: T-major so a T-column is contiguous
  };
  std::vector<Partner> partners;
};

void rb_set_error(const std::string &msg);

// LAMDA molecular-data loader -> SoA tables.
//
// Replaces the parse half of Fortran readdata (emcee/pyradex/radex/radex.so@0x1cf90, called from
// emcee/pyradex/core.py:570,744,887) and the collider discovery pyradex does through astroquery
// (emcee/pyradex/utils.py:53-62).  The reference re-opens and re-parses the file twice per solve
// (core.py:569-570, 741-744); here it is parsed once and the tables live on the device.
//
// Format: blocks separated by '!' comment lines; level rows
// "idx E[cm^-1] g qnum...", line rows "idx up low A[s^-1] freq[GHz] Eup[K]", then per collision
// partner: id line (leading integer 1..7), ncoll, ntemp, temperatures, "idx up low rate(T1..Tn)".
// Like RADEX, the line frequency used for the physics is eterm(up) - eterm(low), not the GHz column.
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "internal.h"

namespace {

thread_local std::string g_err;

// Reads the next line that is not a '!' comment.  RADEX skips comment lines by position; every
// LAMDA file alternates comment/data blocks so skipping by marker accepts the same files.
bool next_data_line(std::istream &in, std::string &line) {
  while (std::getline(in, line)) {
    size_t p = 0;
    while (p < line.size() && std::isspace(static_cast<unsigned char>(line[p]))) ++p;
    if (p == line.size()) continue;
    if (line[p] == '!') continue;
    return true;
  }
  return false;
}

bool fail(const std::string &what, const char *path) {
  rb_set_error(std::string("LAMDA parse error (") + what + ") in " + path);
  return false;
}

bool parse(const char *path, rb_mol &m) {
  std::ifstream in(path);
  if (!in) {
    rb_set_error(std::string("cannot open molecular data file ") + path);
    return false;
  }
  std::string line;
  if (!next_data_line(in, line)) return fail("name", path);
  m.name = line;
  if (!next_data_line(in, line)) return fail("weight", path);
  m.amass = std::strtod(line.c_str(), nullptr);
  if (!next_data_line(in, line)) return fail("nlev", path);
  m.nlev = std::atoi(line.c_str());
  if (m.nlev < 2) return fail("nlev < 2", path);
  m.eterm.resize(m.nlev);
  m.gstat.resize(m.nlev);
  for (int i = 0; i < m.nlev; ++i) {
    if (!next_data_line(in, line)) return fail("level row", path);
    std::istringstream ss(line);
    int idx;
    if (!(ss >> idx >> m.eterm[i] >> m.gstat[i])) return fail("level row", path);
  }
  if (!next_data_line(in, line)) return fail("nline", path);
  m.nline = std::atoi(line.c_str());
  if (m.nline < 1) return fail("nline < 1", path);
  m.iupp.resize(m.nline);
  m.ilow.resize(m.nline);
  m.aeinst.resize(m.nline);
  m.spfreq.resize(m.nline);
  m.eup.resize(m.nline);
  m.xnu.resize(m.nline);
  for (int l = 0; l < m.nline; ++l) {
    if (!next_data_line(in, line)) return fail("line row", path);
    std::istringstream ss(line);
    int idx, up, lo;
    if (!(ss >> idx >> up >> lo >> m.aeinst[l] >> m.spfreq[l] >> m.eup[l])) return fail("line row", path);
    if (up < 1 || up > m.nlev || lo < 1 || lo > m.nlev) return fail("line level index", path);
    m.iupp[l] = up - 1;
    m.ilow[l] = lo - 1;
    m.xnu[l] = m.eterm[up - 1] - m.eterm[lo - 1];
  }
  if (!next_data_line(in, line)) return fail("npart", path);
  m.npart = std::atoi(line.c_str());
  if (m.npart < 1 || m.npart > RB_MAXPART) return fail("npart out of range", path);
  m.partners.resize(m.npart);
  for (int p = 0; p < m.npart; ++p) {
    rb_mol::Partner &pt = m.partners[p];
    if (!next_data_line(in, line)) return fail("partner id", path);
    pt.id = std::atoi(line.c_str());
    if (pt.id < 1 || pt.id > 7) return fail("partner id not in 1..7", path);
    if (!next_data_line(in, line)) return fail("ncoll", path);
    pt.ncoll = std::atoi(line.c_str());
    if (!next_data_line(in, line)) return fail("ntemp", path);
    pt.ntemp = std::atoi(line.c_str());
    if (pt.ncoll < 0 || pt.ntemp < 1) return fail("ncoll/ntemp", path);
    if (!next_data_line(in, line)) return fail("temperatures", path);
    {
      std::istringstream ss(line);
      pt.temps.resize(pt.ntemp);
      for (int t = 0; t < pt.ntemp; ++t)
        if (!(ss >> pt.temps[t])) return fail("temperatures", path);
    }
    pt.lcu.resize(pt.ncoll);
    pt.lcl.resize(pt.ncoll);
    pt.rates_tc.assign(static_cast<size_t>(pt.ntemp) * pt.ncoll, 0.0);
    for (int c = 0; c < pt.ncoll; ++c) {
      if (!next_data_line(in, line)) return fail("collision row", path);
      std::istringstream ss(line);
      int idx, up, lo;
      if (!(ss >> idx >> up >> lo)) return fail("collision row", path);
      if (up < 1 || up > m.nlev || lo < 1 || lo > m.nlev) return fail("collision level index", path);
      pt.lcu[c] = up - 1;
      pt.lcl[c] = lo - 1;
      for (int t = 0; t < pt.ntemp; ++t) {
        double v;
        if (!(ss >> v)) return fail("collision rate", path);
        pt.rates_tc[static_cast<size_t>(t) * pt.ncoll + c] = v;
      }
    }
  }
  return true;
}

}  // namespace

void rb_set_error(const std::string &msg) { g_err = msg; }

extern "C" {

const char *rb_last_error(void) { return g_err.c_str(); }

int rb_moldata_load(const char *path, rb_mol **out) {
  if (!path || !out) {
    rb_set_error("rb_moldata_load: null argument");
    return RB_ERR_ARG;
  }
  rb_mol *m = new rb_mol();
  if (!parse(path, *m)) {
    delete m;
    *out = nullptr;
    return RB_ERR_IO;
  }
  *out = m;
  return RB_OK;
}

void rb_moldata_free(rb_mol *mol) { delete mol; }

int rb_moldata_dims(const rb_mol *mol, int32_t *nlev, int32_t *nline, int32_t *npart) {
  if (!mol) return RB_ERR_ARG;
  if (nlev) *nlev = mol->nlev;
  if (nline) *nline = mol->nline;
  if (npart) *npart = mol->npart;
  return RB_OK;
}

int rb_moldata_partners(const rb_mol *mol, int32_t *partner_id, int32_t *ncoll, int32_t *ntemp) {
  if (!mol) return RB_ERR_ARG;
  for (int p = 0; p < mol->npart; ++p) {
    if (partner_id) partner_id[p] = mol->partners[p].id;
    if (ncoll) ncoll[p] = mol->partners[p].ncoll;
    if (ntemp) ntemp[p] = mol->partners[p].ntemp;
  }
  return RB_OK;
}

int rb_moldata_levels(const rb_mol *mol, double *eterm, double *gstat) {
  if (!mol) return RB_ERR_ARG;
  for (int i = 0; i < mol->nlev; ++i) {
    if (eterm) eterm[i] = mol->eterm[i];
    if (gstat) gstat[i] = mol->gstat[i];
  }
  return RB_OK;
}

int rb_moldata_lines(const rb_mol *mol, int32_t *iupp, int32_t *ilow, double *aeinst, double *spfreq_ghz,
                     double *eup_k, double *xnu) {
  if (!mol) return RB_ERR_ARG;
  for (int l = 0; l < mol->nline; ++l) {
    if (iupp) iupp[l] = mol->iupp[l] + 1;
    if (ilow) ilow[l] = mol->ilow[l] + 1;
    if (aeinst) aeinst[l] = mol->aeinst[l];
    if (spfreq_ghz) spfreq_ghz[l] = mol->spfreq[l];
    if (eup_k) eup_k[l] = mol->eup[l];
    if (xnu) xnu[l] = mol->xnu[l];
  }
  return RB_OK;
}

}  // extern "C"

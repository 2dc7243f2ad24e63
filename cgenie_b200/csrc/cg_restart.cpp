// cg_restart.cpp -- netCDF restart files of GOLDSTEIN, the EMBM and the sea-ice model in the reference's layout
// ("restart/output wire formats"), host code only.
//
// What the reference does (variable names, types, order of definition, masks) is followed routine by routine:
//   outm_netcdf / inm_netcdf   src/goldstein/goldstein_data.f90:153-300 / 11-150
//   outm_netcdf_embm / inm_netcdf_embm   src/embm/embm_data.f90:83-200 / 11-80
//   outm_netcdf_sic / inm_netcdf_sic     src/goldsteinseaice/gold_seaice_data.f90:100-230 / 11-98
// The container format is written by cg_nc3 (no netCDF library in this image).  netCDF-Fortran reverses the order of the
// dimension list, so a Fortran (maxi,maxj,maxk) array is a file variable (depth, latitude, longitude) with the same bytes.
#include "cg_nc3.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/cgenie_b200.h"

namespace cg {

// ------------------------------------------------------------------ CDF-1 codec
namespace {
constexpr uint32_t kDimTag = 0x0A, kVarTag = 0x0B, kAttTag = 0x0C;
int type_size(int t) { return t == NC3_DOUBLE ? 8 : (t == NC3_CHAR ? 1 : 4); }
struct Out {
  std::vector<unsigned char> b;
  void u32(uint32_t v) { for (int s = 24; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
  void u64(uint64_t v) { for (int s = 56; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
  void pad() { while (b.size() % 4) b.push_back(0); }
  void name(const std::string &s) { u32((uint32_t)s.size()); b.insert(b.end(), s.begin(), s.end()); pad(); }
  void atts(const std::vector<cg::Nc3Att> &a) {
    if (a.empty()) { u32(0); u32(0); return; }
    u32(kAttTag); u32((uint32_t)a.size());
    for (auto &t : a) {
      name(t.name); u32((uint32_t)t.type);
      if (t.type == NC3_CHAR) { u32((uint32_t)t.text.size()); b.insert(b.end(), t.text.begin(), t.text.end()); }
      else {
        u32((uint32_t)t.num.size());
        for (double x : t.num) {
          if (t.type == NC3_DOUBLE) { uint64_t u; std::memcpy(&u, &x, 8); u64(u); }
          else if (t.type == NC3_FLOAT) { const float f = (float)x; uint32_t u; std::memcpy(&u, &f, 4); u32(u); }
          else u32((uint32_t)(int32_t)std::llround(x));
        }
      }
      pad();
    }
  }
};
struct In {
  const std::vector<unsigned char> &b; size_t p = 0; bool ok = true;
  explicit In(const std::vector<unsigned char> &b_) : b(b_) {}
  uint32_t u32() { if (p + 4 > b.size()) { ok = false; return 0; } uint32_t v = 0; for (int q = 0; q < 4; q++) v = (v << 8) | b[p++]; return v; }
  uint64_t u64() { uint64_t hi = u32(); return (hi << 32) | u32(); }
  std::string name() {
    const uint32_t n = u32();
    if (!ok || p + n > b.size()) { ok = false; return ""; }
    std::string s((const char *)&b[p], n);
    p += (n + 3) / 4 * 4;
    return s;
  }
  void atts(std::vector<cg::Nc3Att> *out) {
    const uint32_t tag = u32(), n = u32();
    if (tag == 0 && n == 0) return;
    if (tag != kAttTag) { ok = false; return; }
    for (uint32_t q = 0; q < n && ok; q++) {
      cg::Nc3Att t;
      t.name = name();
      t.type = (int)u32();
      const uint32_t ne = u32();
      const size_t es = (t.type == NC3_CHAR || t.type == 1) ? 1 : t.type == 3 ? 2 : t.type == NC3_DOUBLE ? 8 : 4;
      const size_t bytes = (size_t)ne * es;
      if (p + bytes > b.size()) { ok = false; return; }
      if (t.type == NC3_CHAR) t.text.assign((const char *)&b[p], ne);
      else if (t.type == NC3_DOUBLE || t.type == NC3_FLOAT || t.type == NC3_INT)
        for (uint32_t e = 0; e < ne; e++) {
          const unsigned char *c = &b[p + e * es];
          uint64_t u = 0;
          for (size_t s2 = 0; s2 < es; s2++) u = (u << 8) | c[s2];
          if (t.type == NC3_DOUBLE) { double d; std::memcpy(&d, &u, 8); t.num.push_back(d); }
          else if (t.type == NC3_FLOAT) { const uint32_t u4 = (uint32_t)u; float f; std::memcpy(&f, &u4, 4); t.num.push_back(f); }
          else t.num.push_back((int32_t)(uint32_t)u);
        }
      if (out) out->push_back(t);
      p += (bytes + 3) / 4 * 4;
    }
  }
};
}  // namespace

int Nc3File::add_dim(const std::string &name, int len) { dims.push_back({name, len}); return (int)dims.size() - 1; }
int Nc3File::add_var(const std::string &name, int type, const std::vector<int> &dimids) {
  Nc3Var v;
  v.name = name; v.type = type; v.dimids = dimids; v.count = 1;
  v.rec = !dimids.empty() && dims[dimids[0]].second == 0;
  for (size_t q = v.rec ? 1 : 0; q < dimids.size(); q++) v.count *= dims[dimids[q]].second;
  v.data.assign((size_t)v.count * (size_t)(v.rec ? numrecs : 1), 0.0);
  vars.push_back(v);
  return (int)vars.size() - 1;
}
void Nc3File::put_att(int varid, const std::string &name, const std::string &value) {
  Nc3Att t; t.name = name; t.type = NC3_CHAR; t.text = value;
  (varid < 0 ? gatts : vars[varid].atts).push_back(t);
}
void Nc3File::put_att_num(int varid, const std::string &name, int type, const std::vector<double> &v) {
  Nc3Att t; t.name = name; t.type = type; t.num = v;
  (varid < 0 ? gatts : vars[varid].atts).push_back(t);
}
void Nc3File::put(int varid, const double *v, long long n) { for (long long q = 0; q < n && q < vars[varid].count; q++) vars[varid].data[q] = v[q]; }
void Nc3File::put(int varid, const int *v, long long n) { for (long long q = 0; q < n && q < vars[varid].count; q++) vars[varid].data[q] = v[q]; }
void Nc3File::put_rec(int varid, int rec, const double *v, long long n) {
  if (rec + 1 > numrecs) numrecs = rec + 1;
  for (auto &w : vars) if (w.rec && w.data.size() < (size_t)w.count * (size_t)numrecs) w.data.resize((size_t)w.count * (size_t)numrecs, 0.0);
  Nc3Var &w = vars[varid];
  for (long long q = 0; q < n && q < w.count; q++) w.data[(size_t)rec * (size_t)w.count + (size_t)q] = v[q];
}
const Nc3Var *Nc3File::var(const std::string &name) const {
  for (auto &v : vars) if (v.name == name) return &v;
  return nullptr;
}
int Nc3File::dim_len(const std::string &name) const {
  for (auto &d : dims) if (d.first == name) return d.second;
  return -1;
}

bool Nc3File::write(const std::string &path, std::string *err) const {
  auto vsize = [](const Nc3Var &v) { return ((size_t)v.count * type_size(v.type) + 3) / 4 * 4; };   // record variables: of one record
  auto header = [&](const std::vector<uint32_t> &begin) {
    Out o;
    o.b = {'C', 'D', 'F', 1};
    o.u32((uint32_t)numrecs);
    if (dims.empty()) { o.u32(0); o.u32(0); }
    else { o.u32(kDimTag); o.u32((uint32_t)dims.size()); for (auto &d : dims) { o.name(d.first); o.u32((uint32_t)d.second); } }
    o.atts(gatts);
    if (vars.empty()) { o.u32(0); o.u32(0); }
    else {
      o.u32(kVarTag); o.u32((uint32_t)vars.size());
      for (size_t q = 0; q < vars.size(); q++) {
        const Nc3Var &v = vars[q];
        o.name(v.name);
        o.u32((uint32_t)v.dimids.size());
        for (int d : v.dimids) o.u32((uint32_t)d);
        o.atts(v.atts);
        o.u32((uint32_t)v.type);
        o.u32((uint32_t)vsize(v));
        o.u32(begin[q]);
      }
    }
    return o;
  };
  // the fixed-size variables in definition order, then the records: per record one slab of every record variable, in definition order
  std::vector<uint32_t> begin(vars.size(), 0);
  size_t off = header(begin).b.size(), recsize = 0;
  for (int pass = 0; pass < 2; pass++)
    for (size_t q = 0; q < vars.size(); q++) {
      if (vars[q].rec != (pass == 1)) continue;
      begin[q] = (uint32_t)off;
      off += vsize(vars[q]);
      if (pass) recsize += vsize(vars[q]);
    }
  if (off + recsize * (size_t)(numrecs > 0 ? numrecs - 1 : 0) > 0x7fffffffULL) { if (err) *err = "file too large for the classic format"; return false; }
  Out o = header(begin);
  auto emit = [&](const Nc3Var &v, size_t q0) {
    if (v.data.size() < q0 + (size_t)v.count) { o.b.insert(o.b.end(), vsize(v), 0); return; }
    for (size_t q = q0; q < q0 + (size_t)v.count; q++) {
      if (v.type == NC3_DOUBLE) { uint64_t u; std::memcpy(&u, &v.data[q], 8); o.u64(u); }
      else if (v.type == NC3_FLOAT) { const float f = (float)v.data[q]; uint32_t u; std::memcpy(&u, &f, 4); o.u32(u); }
      else if (v.type == NC3_INT) o.u32((uint32_t)(int32_t)std::llround(v.data[q]));
      else o.b.push_back((unsigned char)v.data[q]);
    }
    o.pad();
  };
  for (auto &v : vars) if (!v.rec) emit(v, 0);
  for (int r = 0; r < numrecs; r++)
    for (auto &v : vars) if (v.rec) emit(v, (size_t)r * (size_t)v.count);
  FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) { if (err) *err = "cannot open " + path + " for writing"; return false; }
  const bool ok = std::fwrite(o.b.data(), 1, o.b.size(), f) == o.b.size();
  std::fclose(f);
  if (!ok && err) *err = "short write to " + path;
  return ok;
}

bool Nc3File::read(const std::string &path, std::string *err) {
  dims.clear(); gatts.clear(); vars.clear(); numrecs = 0;
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) { if (err) *err = "Missing file " + path; return false; }
  std::vector<unsigned char> b;
  unsigned char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) b.insert(b.end(), buf, buf + n);
  std::fclose(f);
  if (b.size() < 8 || b[0] != 'C' || b[1] != 'D' || b[2] != 'F' || (b[3] != 1 && b[3] != 2)) {
    if (err) *err = path + ": not a netCDF classic (CDF-1/2) file";
    return false;
  }
  const bool wide = b[3] == 2;
  In in(b);
  in.p = 4;
  numrecs = (int)in.u32();
  if (numrecs < 0) numrecs = 0;   // NC_STREAMING (0xffffffff) is never written for these files
  int recdim = -1;
  {
    const uint32_t tag = in.u32(), nd = in.u32();
    if (!(tag == 0 && nd == 0)) {
      if (tag != kDimTag) in.ok = false;
      for (uint32_t q = 0; q < nd && in.ok; q++) {
        const std::string nm = in.name();
        const int len = (int)in.u32();
        if (len == 0) recdim = (int)q;
        dims.push_back({nm, len});
      }
    }
  }
  in.atts(&gatts);
  std::vector<uint64_t> begin;
  std::vector<uint32_t> vsz;
  {
    const uint32_t tag = in.u32(), nv = in.u32();
    if (!(tag == 0 && nv == 0)) {
      if (tag != kVarTag) in.ok = false;
      for (uint32_t q = 0; q < nv && in.ok; q++) {
        Nc3Var v;
        v.name = in.name();
        const uint32_t nd = in.u32();
        v.count = 1;
        for (uint32_t d = 0; d < nd && in.ok; d++) {
          const int id = (int)in.u32();
          if (id < 0 || id >= (int)dims.size()) { in.ok = false; break; }
          if (id == recdim) {
            if (d != 0) { if (err) *err = path + ": record dimension not the first of variable " + v.name; return false; }
            v.rec = true;
          }
          v.dimids.push_back(id);
          if (id != recdim) v.count *= dims[id].second;
        }
        in.atts(&v.atts);
        v.type = (int)in.u32();
        vsz.push_back(in.u32());
        begin.push_back(wide ? in.u64() : in.u32());
        vars.push_back(v);
      }
    }
  }
  if (!in.ok) { if (err) *err = path + ": damaged netCDF header"; return false; }
  // one record = the slabs of all record variables (vsize each; a lone record variable is not padded, which the 4- and 8-byte
  // types of these files never need)
  size_t recsize = 0;
  for (size_t q = 0; q < vars.size(); q++) if (vars[q].rec) recsize += vsz[q];
  for (size_t q = 0; q < vars.size(); q++) {
    Nc3Var &v = vars[q];
    if (v.type != NC3_DOUBLE && v.type != NC3_FLOAT && v.type != NC3_INT && v.type != NC3_CHAR) {
      if (err) *err = path + ": unsupported type of variable " + v.name;
      return false;
    }
    const size_t ts = (size_t)type_size(v.type);
    const size_t nrec = v.rec ? (size_t)numrecs : 1;
    if (nrec && begin[q] + (uint64_t)(nrec - 1) * recsize + (uint64_t)v.count * ts > b.size()) { if (err) *err = path + ": truncated data of variable " + v.name; return false; }
    v.data.resize((size_t)v.count * nrec);
    for (size_t r = 0; r < nrec; r++) {
      const unsigned char *p = &b[begin[q] + r * recsize];
      double *out = &v.data[r * (size_t)v.count];
      for (long long e = 0; e < v.count; e++, p += ts) {
        if (v.type == NC3_DOUBLE) { uint64_t u = 0; for (int s = 0; s < 8; s++) u = (u << 8) | p[s]; double d; std::memcpy(&d, &u, 8); out[e] = d; }
        else if (v.type == NC3_CHAR) out[e] = p[0];
        else {
          uint32_t u = 0; for (int s = 0; s < 4; s++) u = (u << 8) | p[s];
          if (v.type == NC3_FLOAT) { float x; std::memcpy(&x, &u, 4); out[e] = x; } else out[e] = (int32_t)u;
        }
      }
    }
  }
  return true;
}

}  // namespace cg

// ------------------------------------------------------------------ restart layouts (C ABI, plain host arrays)
namespace {
thread_local std::string g_restart_err;
int rfail(const std::string &m) { g_restart_err = m; return CG_ERR_IO; }

// dimensions and the scalar date variables every module's restart starts with
struct Common { int nrecs, lon, lat, dep = -1, v_lon, v_lat, v_dep = -1; };
Common define_axes(cg::Nc3File &f, int maxi, int maxj, int maxk, bool axis_atts) {
  Common c;
  c.nrecs = f.add_dim("nrecs", 1);
  c.lon = f.add_dim("longitude", maxi);
  c.lat = f.add_dim("latitude", maxj);
  if (maxk > 0) c.dep = f.add_dim("depth", maxk);
  c.v_lon = f.add_var("longitude", cg::NC3_FLOAT, {c.lon});
  c.v_lat = f.add_var("latitude", cg::NC3_FLOAT, {c.lat});
  if (maxk > 0) c.v_dep = f.add_var("depth", cg::NC3_FLOAT, {c.dep});
  if (axis_atts) {   // goldstein_data.f90:271-274
    f.put_att(c.v_lon, "units", "degrees_east"); f.put_att(c.v_lon, "long_name", "longitude");
    f.put_att(c.v_lat, "units", "degrees_north"); f.put_att(c.v_lat, "long_name", "latitude");
  }
  return c;
}
void define_date(cg::Nc3File &f, const Common &c, const int32_t date[4]) {
  // definition order ioffset, iyear, imonth, iday (goldstein_data.f90:247-251); date = {iyear, imonth, iday, ioffset}
  const int order[4] = {3, 0, 1, 2};
  const char *names[4] = {"ioffset", "iyear", "imonth", "iday"};
  for (int q = 0; q < 4; q++) { const int id = f.add_var(names[q], cg::NC3_INT, {c.nrecs}); const int v = date[order[q]]; f.put(id, &v, 1); }
}
bool get_date(const cg::Nc3File &f, int32_t date[4]) {
  const char *names[4] = {"iyear", "imonth", "iday", "ioffset"};
  for (int q = 0; q < 4; q++) {
    const cg::Nc3Var *v = f.var(names[q]);
    if (!v || v->count < 1) return false;
    if (date) date[q] = (int32_t)v->data[0];
  }
  return true;
}
const cg::Nc3Var *need(const cg::Nc3File &f, const char *name, long long count, std::string *err) {
  const cg::Nc3Var *v = f.var(name);
  if (!v) { *err = std::string("variable ") + name + " missing"; return nullptr; }
  if (v->count != count) { *err = std::string("variable ") + name + " has the wrong size"; return nullptr; }
  return v;
}
}  // namespace

extern "C" const char *cg_restart_last_error(void) { return g_restart_err.c_str(); }

// outm_netcdf, goldstein_data.f90:153-300.  ts (maxl,maxi,maxj,maxk), u (3,maxi,maxj,maxk), k1 (0:maxi+1,0:maxj+1), all Fortran order.
extern "C" int cg_restart_goldstein_write(const char *path, int maxi, int maxj, int maxk, int maxl, const int32_t *k1,
                                          const double *lon, const double *lat, const double *depth, const double *ts,
                                          const double *u, const double *evap, const double *late, const double *sens,
                                          const int32_t date[4]) {
  if (!path || !k1 || !lon || !lat || !depth || !ts || !u || !date || maxl < 2) return rfail("cg_restart_goldstein_write: bad argument");
  cg::Nc3File f;
  const Common c = define_axes(f, maxi, maxj, maxk, true);
  define_date(f, c, date);
  const size_t n3 = (size_t)maxi * maxj * maxk, n2 = (size_t)maxi * maxj;
  std::vector<double> a(n3);
  const char *n3d[4] = {"temp", "salinity", "uvel", "vvel"};
  int id3[4], id2[3];
  for (int q = 0; q < 4; q++) id3[q] = f.add_var(n3d[q], cg::NC3_DOUBLE, {c.dep, c.lat, c.lon});
  const char *n2d[3] = {"evap", "late", "sens"};
  for (int q = 0; q < 3; q++) id2[q] = f.add_var(n2d[q], cg::NC3_DOUBLE, {c.lat, c.lon});
  f.put(c.v_lon, lon, maxi); f.put(c.v_lat, lat, maxj); f.put(c.v_dep, depth, maxk);
  for (int q = 0; q < 4; q++) {
    for (int k = 0; k < maxk; k++)
      for (int j = 0; j < maxj; j++)
        for (int i = 0; i < maxi; i++) {
          const size_t cell = (size_t)i + (size_t)maxi * (j + (size_t)maxj * k);
          // landmask(i,j,:) = 1 wherever ANY level of the column is wet (goldstein_data.f90:200-207): ocean columns are
          // written whole, land columns as zeros; velocities are not masked (:212-213)
          const bool ocean = k1[(i + 1) + (size_t)(maxi + 2) * (j + 1)] <= maxk;
          a[cell] = q < 2 ? (ocean ? ts[q + (size_t)maxl * cell] : 0.0) : u[(q - 2) + 3 * cell];
        }
    f.put(id3[q], a.data(), (long long)n3);
  }
  const double *two[3] = {evap, late, sens};
  for (int q = 0; q < 3; q++) if (two[q]) f.put(id2[q], two[q], (long long)n2);
  std::string err;
  if (!f.write(path, &err)) return rfail(err);
  return CG_OK;
}
// inm_netcdf, goldstein_data.f90:11-150: T and S replace ts(1:2), uvel/vvel replace u(1:2) (the caller copies u to u1,
// :94-96); evap/late/sens only if wanted (lrestart_genie).  date = the file's {iyear, imonth, iday, ioffset}.
extern "C" int cg_restart_goldstein_read(const char *path, int maxi, int maxj, int maxk, int maxl, double *ts, double *u,
                                         double *evap, double *late, double *sens, int32_t date[4]) {
  if (!path || !ts || !u || maxl < 2) return rfail("cg_restart_goldstein_read: bad argument");
  cg::Nc3File f;
  std::string err;
  if (!f.read(path, &err)) return rfail(err);
  if (!get_date(f, date)) return rfail(std::string(path) + ": date variables missing");
  const long long n3 = (long long)maxi * maxj * maxk, n2 = (long long)maxi * maxj;
  const char *n3d[4] = {"temp", "salinity", "uvel", "vvel"};
  for (int q = 0; q < 4; q++) {
    const cg::Nc3Var *v = need(f, n3d[q], n3, &err);
    if (!v) return rfail(std::string(path) + ": " + err);
    for (long long cell = 0; cell < n3; cell++) {
      if (q < 2) ts[q + (size_t)maxl * cell] = v->data[cell]; else u[(q - 2) + 3 * cell] = v->data[cell];
    }
  }
  const char *n2d[3] = {"evap", "late", "sens"};
  double *two[3] = {evap, late, sens};
  for (int q = 0; q < 3; q++) {
    if (!two[q]) continue;
    const cg::Nc3Var *v = need(f, n2d[q], n2, &err);
    if (!v) return rfail(std::string(path) + ": " + err);
    std::memcpy(two[q], v->data.data(), (size_t)n2 * 8);
  }
  return CG_OK;
}

// outm_netcdf_embm, embm_data.f90:83-200: tq (2,maxi,maxj) -> air_temp, humidity
extern "C" int cg_restart_embm_write(const char *path, int maxi, int maxj, const double *lon, const double *lat, const double *tq,
                                     const int32_t date[4]) {
  if (!path || !lon || !lat || !tq || !date) return rfail("cg_restart_embm_write: bad argument");
  cg::Nc3File f;
  const Common c = define_axes(f, maxi, maxj, 0, false);
  define_date(f, c, date);
  const size_t n2 = (size_t)maxi * maxj;
  const int idt = f.add_var("air_temp", cg::NC3_DOUBLE, {c.lat, c.lon}), idq = f.add_var("humidity", cg::NC3_DOUBLE, {c.lat, c.lon});
  f.put(c.v_lon, lon, maxi); f.put(c.v_lat, lat, maxj);
  std::vector<double> a(n2), b(n2);
  for (size_t q = 0; q < n2; q++) { a[q] = tq[2 * q]; b[q] = tq[2 * q + 1]; }
  f.put(idt, a.data(), (long long)n2); f.put(idq, b.data(), (long long)n2);
  std::string err;
  if (!f.write(path, &err)) return rfail(err);
  return CG_OK;
}
// inm_netcdf_embm, embm_data.f90:11-80
extern "C" int cg_restart_embm_read(const char *path, int maxi, int maxj, double *tq, int32_t date[4]) {
  if (!path || !tq) return rfail("cg_restart_embm_read: bad argument");
  cg::Nc3File f;
  std::string err;
  if (!f.read(path, &err)) return rfail(err);
  if (!get_date(f, date)) return rfail(std::string(path) + ": date variables missing");
  const long long n2 = (long long)maxi * maxj;
  const cg::Nc3Var *t = need(f, "air_temp", n2, &err), *q = t ? need(f, "humidity", n2, &err) : nullptr;
  if (!t || !q) return rfail(std::string(path) + ": " + err);
  for (long long c = 0; c < n2; c++) { tq[2 * c] = t->data[c]; tq[2 * c + 1] = q->data[c]; }
  return CG_OK;
}

// outm_netcdf_sic, gold_seaice_data.f90:100-230: varice (2,maxi,maxj) masked by k1 < 90 (:141-148), tice, albice unmasked
extern "C" int cg_restart_seaice_write(const char *path, int maxi, int maxj, const int32_t *k1, const double *lon, const double *lat,
                                       const double *varice, const double *tice, const double *albice, const int32_t date[4]) {
  if (!path || !k1 || !lon || !lat || !varice || !tice || !albice || !date) return rfail("cg_restart_seaice_write: bad argument");
  cg::Nc3File f;
  const Common c = define_axes(f, maxi, maxj, 0, false);
  define_date(f, c, date);
  const size_t n2 = (size_t)maxi * maxj;
  const char *names[4] = {"sic_height", "sic_cover", "sic_temp", "sic_albedo"};
  int id[4];
  for (int q = 0; q < 4; q++) id[q] = f.add_var(names[q], cg::NC3_DOUBLE, {c.lat, c.lon});
  f.put(c.v_lon, lon, maxi); f.put(c.v_lat, lat, maxj);
  std::vector<double> a(n2), b(n2);
  for (int j = 0; j < maxj; j++)
    for (int i = 0; i < maxi; i++) {
      const size_t q = (size_t)i + (size_t)maxi * j;
      const bool ocean = k1[(i + 1) + (size_t)(maxi + 2) * (j + 1)] < 90;
      a[q] = ocean ? varice[2 * q] : 0.0;
      b[q] = ocean ? varice[2 * q + 1] : 0.0;
    }
  f.put(id[0], a.data(), (long long)n2); f.put(id[1], b.data(), (long long)n2);
  f.put(id[2], tice, (long long)n2); f.put(id[3], albice, (long long)n2);
  std::string err;
  if (!f.write(path, &err)) return rfail(err);
  return CG_OK;
}
// inm_netcdf_sic, gold_seaice_data.f90:11-98
extern "C" int cg_restart_seaice_read(const char *path, int maxi, int maxj, double *varice, double *tice, double *albice, int32_t date[4]) {
  if (!path || !varice || !tice || !albice) return rfail("cg_restart_seaice_read: bad argument");
  cg::Nc3File f;
  std::string err;
  if (!f.read(path, &err)) return rfail(err);
  if (!get_date(f, date)) return rfail(std::string(path) + ": date variables missing");
  const long long n2 = (long long)maxi * maxj;
  const char *names[4] = {"sic_height", "sic_cover", "sic_temp", "sic_albedo"};
  const cg::Nc3Var *v[4];
  for (int q = 0; q < 4; q++) { v[q] = need(f, names[q], n2, &err); if (!v[q]) return rfail(std::string(path) + ": " + err); }
  for (long long c = 0; c < n2; c++) { varice[2 * c] = v[0]->data[c]; varice[2 * c + 1] = v[1]->data[c]; tice[c] = v[2]->data[c]; albice[c] = v[3]->data[c]; }
  return CG_OK;
}
// the date block alone (genie-main's main_restart_N.nc: data/main/main_restart_0.nc is one, written by the netCDF library)
extern "C" int cg_restart_date_write(const char *path, const int32_t date[4]) {
  if (!path || !date) return rfail("cg_restart_date_write: bad argument");
  cg::Nc3File f;
  Common c;
  c.nrecs = f.add_dim("nrecs", 1);
  define_date(f, c, date);
  std::string err;
  if (!f.write(path, &err)) return rfail(err);
  return CG_OK;
}
extern "C" int cg_restart_date_read(const char *path, int32_t date[4]) {
  if (!path || !date) return rfail("cg_restart_date_read: bad argument");
  cg::Nc3File f;
  std::string err;
  if (!f.read(path, &err)) return rfail(err);
  if (f.dim_len("nrecs") != 1 || !get_date(f, date)) return rfail(std::string(path) + ": date variables missing");
  return CG_OK;
}

// ------------------------------------------------------------------ BIOGEM restart (ctrl_ncrst = .TRUE., the default)
// sub_data_netCDF_ncrstsave, src/biogem/biogem_data_netCDF.f90:24-142, through the helpers of src/common/gem_netcdf.f90
// (sub_putglobal :176-217, sub_defvar :250-359, sub_putvar1d :569-602, sub_putvar3d :746-786, edge_maker :877-913).
namespace {
constexpr double kNcFillDouble = 9.9692099683868690e+36;   // nf90_fill_double
// sub_defvar: [valid_range only if rmin != rmax -- never here], missing_value (always a DOUBLE attribute, also on float
// variables), axis, edges (axis variables that are not themselves *_edges), long_name, standard_name, units
int defvar(cg::Nc3File &f, const std::string &name, int type, const std::vector<int> &dimids, const std::string &axis,
           const std::string &lname, const std::string &sname, const std::string &units) {
  const int id = f.add_var(name, type, dimids);
  f.put_att_num(id, "missing_value", cg::NC3_DOUBLE, {kNcFillDouble});
  if (axis != " ") {
    f.put_att(id, "axis", axis);
    const bool is_edges = name.size() >= 6 && name.compare(name.size() - 6, 6, "_edges") == 0;
    if (axis != "T" && !is_edges) f.put_att(id, "edges", name + "_edges");
  }
  if (lname != " ") f.put_att(id, "long_name", lname);
  if (sname != " ") f.put_att(id, "standard_name", sname);
  if (units != " ") f.put_att(id, "units", units);
  return id;
}
}  // namespace

// ocn (n_ocn,n_i,n_j,n_k) and bio_part (n_sed,n_i,n_j,n_k) in Fortran order, k = 1 the deepest level; the file holds every
// tracer as a FLOAT variable (zt, lat, lon) with the surface first and the fill value on dry cells (mask = k >= k1(i,j)).
// lon / lat (t grid), lon_e (0:n_i), lat_e (0:n_j), zt (n_k, surface first) and zt_e (0:n_k) are the axes the reference
// takes from phys_ocn (biogem_data.f90:1115-1123) through edge_maker.
extern "C" int cg_restart_biogem_write(const char *path, int n_i, int n_j, int n_k, const int32_t *k1, const double *lon,
                                       const double *lat, const double *lon_e, const double *lat_e, const double *zt,
                                       const double *zt_e, int n_ocn, const char *const *ocn_names,
                                       const char *const *ocn_longnames, const double *ocn, int n_sed,
                                       const char *const *sed_names, const char *const *sed_longnames, const double *bio_part,
                                       double year, const char *run_id) {
  if (!path || !k1 || !lon || !lat || !lon_e || !lat_e || !zt || !zt_e || n_ocn < 0 || n_sed < 0 || (n_ocn && (!ocn_names || !ocn_longnames || !ocn)) ||
      (n_sed && (!sed_names || !sed_longnames || !bio_part)))
    return rfail("cg_restart_biogem_write: bad argument");
  cg::Nc3File f;
  // sub_putglobal; loc_string_year is CHARACTER(7) and receives the 8-digit fun_conv_num_char_n(8, int(yr)): the last
  // digit is cut off (biogem_data_netCDF.f90:41,62).  loc_timunit is never set there; it is left out.
  char y8[16];
  std::snprintf(y8, sizeof y8, "%08d", (int)year);
  f.put_att(-1, "Conventions", "CF-1.0");
  f.put_att(-1, "file_name", path);
  f.put_att(-1, "title", std::string("BIOGEM restart @ year ") + std::string(y8).substr(0, 7));
  if (run_id && *run_id) f.put_att(-1, "experiment_name", run_id);
  const int d_lon = f.add_dim("lon", n_i), d_lat = f.add_dim("lat", n_j), d_lone = f.add_dim("lon_edges", n_i + 1),
            d_late = f.add_dim("lat_edges", n_j + 1), d_zt = f.add_dim("zt", n_k), d_zte = f.add_dim("zt_edges", n_k + 1);
  const int v_lon = defvar(f, "lon", cg::NC3_DOUBLE, {d_lon}, "X", "longitude of the t grid", "longitude", "degrees_east");
  const int v_lat = defvar(f, "lat", cg::NC3_DOUBLE, {d_lat}, "Y", "latitude of the t grid", "latitude", "degrees_north");
  const int v_lone = defvar(f, "lon_edges", cg::NC3_DOUBLE, {d_lone}, " ", "longitude of t grid edges", " ", "degrees");
  const int v_late = defvar(f, "lat_edges", cg::NC3_DOUBLE, {d_late}, " ", "latitude of t grid edges", " ", "degrees");
  const int v_zt = defvar(f, "zt", cg::NC3_DOUBLE, {d_zt}, "Z", "depth of z grid", " ", "cm");
  const int v_zte = defvar(f, "zt_edges", cg::NC3_DOUBLE, {d_zte}, " ", "depth of z grid edges", " ", "m");
  std::vector<int> ido(n_ocn), ids(n_sed);
  for (int l = 0; l < n_ocn; l++)
    ido[l] = defvar(f, std::string("ocn_") + ocn_names[l], cg::NC3_FLOAT, {d_zt, d_lat, d_lon}, " ", ocn_longnames[l],
                    std::string("Ocean tracer - ") + ocn_names[l], " ");
  for (int l = 0; l < n_sed; l++)
    ids[l] = defvar(f, std::string("bio_part_") + sed_names[l], cg::NC3_FLOAT, {d_zt, d_lat, d_lon}, " ", sed_longnames[l],
                    std::string("Particulate tracer - ") + sed_names[l], " ");
  f.put(v_lon, lon, n_i); f.put(v_lat, lat, n_j); f.put(v_lone, lon_e, n_i + 1); f.put(v_late, lat_e, n_j + 1);
  f.put(v_zt, zt, n_k); f.put(v_zte, zt_e, n_k + 1);
  const size_t n3 = (size_t)n_i * n_j * n_k;
  std::vector<double> a(n3);
  for (int pass = 0; pass < 2; pass++) {
    const int nt = pass ? n_sed : n_ocn;
    const double *src = pass ? bio_part : ocn;
    for (int l = 0; l < nt; l++) {
      for (int k = 0; k < n_k; k++)       // file level k (0 = surface) = model level n_k - k (loc_ijk(:,:,n_k:1:-1))
        for (int j = 0; j < n_j; j++)
          for (int i = 0; i < n_i; i++) {
            const int km = n_k - k;       // 1-based model level
            const bool wet = km >= k1[(i + 1) + (size_t)(n_i + 2) * (j + 1)];
            const size_t cell = (size_t)i + (size_t)n_i * (j + (size_t)n_j * (km - 1));
            a[(size_t)i + (size_t)n_i * (j + (size_t)n_j * k)] = wet ? src[l + (size_t)nt * cell] : kNcFillDouble;
          }
      f.put(pass ? ids[l] : ido[l], a.data(), (long long)n3);
    }
  }
  std::string err;
  if (!f.write(path, &err)) return rfail(err);
  return CG_OK;
}
// sub_data_load_rst, biogem_data.f90:438-568 (netCDF branch) with sub_getvarijk (gem_netcdf.f90:1120-1151): every selected
// tracer whose variable is in the file is replaced (levels flipped back); tracers without a variable keep their values.
// The reference also stores the file's fill value on dry cells; here dry cells are left alone (nothing reads them).
// found_ocn / found_sed (may be NULL): 1 for tracers that were in the file.
extern "C" int cg_restart_biogem_read(const char *path, int n_i, int n_j, int n_k, const int32_t *k1, int n_ocn,
                                      const char *const *ocn_names, double *ocn, int32_t *found_ocn, int n_sed,
                                      const char *const *sed_names, double *bio_part, int32_t *found_sed) {
  if (!path || !k1 || (n_ocn && (!ocn_names || !ocn)) || (n_sed && (!sed_names || !bio_part))) return rfail("cg_restart_biogem_read: bad argument");
  cg::Nc3File f;
  std::string err;
  if (!f.read(path, &err)) return rfail(err);
  const long long n3 = (long long)n_i * n_j * n_k;
  for (int pass = 0; pass < 2; pass++) {
    const int nt = pass ? n_sed : n_ocn;
    double *dst = pass ? bio_part : ocn;
    for (int l = 0; l < nt; l++) {
      const std::string name = std::string(pass ? "bio_part_" : "ocn_") + (pass ? sed_names : ocn_names)[l];
      const cg::Nc3Var *v = f.var(name);
      int32_t *found = pass ? found_sed : found_ocn;
      if (found) found[l] = v ? 1 : 0;
      if (!v) continue;
      if (v->count != n3) return rfail(std::string(path) + ": variable " + name + " has the wrong size");
      for (int k = 0; k < n_k; k++)
        for (int j = 0; j < n_j; j++)
          for (int i = 0; i < n_i; i++) {
            const int km = n_k - k;
            if (km < k1[(i + 1) + (size_t)(n_i + 2) * (j + 1)]) continue;
            const size_t cell = (size_t)i + (size_t)n_i * (j + (size_t)n_j * (km - 1));
            dst[l + (size_t)nt * cell] = v->data[(size_t)i + (size_t)n_i * (j + (size_t)n_j * k)];
          }
    }
  }
  return CG_OK;
}

// ------------------------------------------------------------------ BIOGEM time slices: fields_biogem_3d.nc
// sub_init_netcdf (dd = 3), sub_save_netcdf, sub_save_netcdf_3d: src/biogem/biogem_data_netCDF.f90:148-277, 282-459, 1959-2315
namespace {
constexpr double kSliceNullSmall = 0.999999e-19;   // const_real_nullsmall, gem_cmn.f90:719
constexpr double kSliceNull = -0.999999e+19;       // const_real_null, :717 (the null the time-slice writer passes, not const_nulliso)
constexpr double kStd13C = 0.011202, kStd14C = 1.176e-12;   // const_standards(11:12), gem_cmn.f90:629-631 (as cg_series.cpp)
// fun_calc_isotope_delta(tot, iso, standard, .FALSE., const_real_null), gem_util.f90:568-598
double slice_delta(double tot, double iso, double standard) {
  if (tot > kSliceNullSmall) {
    const double f = iso / tot;
    if ((1.0 - f) > kSliceNullSmall) {
      const double R = f / (1.0 - f);
      return 1000.0 * (R / standard - 1.0);
    }
  }
  return kSliceNull;
}
double slice_standard(int type) { return type == 11 ? kStd13C : kStd14C; }
// sub_adddef_netcdf (dino = 4) + sub_defvar ('F'): [valid_range as two floats if min != max,] missing_value, long_name, units
int def_field(cg::Nc3File &f, const std::string &name, const std::vector<int> &dimids, const std::string &lname,
              const std::string &units, double rmin, double rmax) {
  if (f.var(name)) { for (size_t q = 0; q < f.vars.size(); q++) if (f.vars[q].name == name) return (int)q; }
  const int id = f.add_var(name, cg::NC3_FLOAT, dimids);
  if (rmin != rmax) f.put_att_num(id, "valid_range", cg::NC3_FLOAT, {rmin, rmax});
  f.put_att_num(id, "missing_value", cg::NC3_DOUBLE, {kNcFillDouble});
  if (lname != " " && !lname.empty()) f.put_att(id, "long_name", lname);
  if (units != " " && !units.empty()) f.put_att(id, "units", units);
  return id;
}
}  // namespace

extern "C" int cg_slice_biogem_write_3d(const char *path, int n_i, int n_j, int n_k, const int32_t *k1, const double *lon,
                                        const double *lat, const double *lon_e, const double *lat_e, const double *zt,
                                        const double *zt_e, int n_ocn, const char *const *ocn_names,
                                        const char *const *ocn_longnames, const char *const *ocn_units, const double *ocn_mima,
                                        const int32_t *ocn_type, const int32_t *ocn_dep, const double *int_ocn, int n_sed,
                                        const char *const *sed_names, const int32_t *sed_type, const int32_t *sed_dep,
                                        const double *int_part, int n_carb, const char *const *carb_names, const double *int_carb,
                                        int n_carbconst, const char *const *carbconst_names, const double *int_carbconst,
                                        const double *mass, double int_t, double year_mid, const char *run_id) {
  if (!path || n_i <= 0 || n_j <= 0 || n_k <= 0 || !k1 || !lon || !lat || !lon_e || !lat_e || !zt || !zt_e || n_ocn < 0 || n_sed < 0 ||
      n_carb < 0 || n_carbconst < 0 ||
      (n_ocn && (!ocn_names || !ocn_longnames || !ocn_units || !ocn_mima || !ocn_type || !ocn_dep || !int_ocn)) ||
      (n_sed && (!sed_names || !sed_type || !sed_dep || !int_part)) || (n_carb && (!carb_names || !int_carb)) ||
      (n_carbconst && (!carbconst_names || !int_carbconst)))
    return rfail("cg_slice_biogem_write_3d: bad argument");
  if (!(int_t > 0.0)) return rfail("cg_slice_biogem_write_3d: int_t_timeslice is not positive (no step of the save window was integrated)");
  for (int l = 0; l < n_ocn; l++) if (ocn_dep[l] < 0 || ocn_dep[l] >= n_ocn) return rfail("cg_slice_biogem_write_3d: ocn_dep out of range");
  for (int l = 0; l < n_sed; l++) if (sed_dep[l] < 0 || sed_dep[l] >= n_sed) return rfail("cg_slice_biogem_write_3d: sed_dep out of range");
  cg::Nc3File f;
  std::string err;
  int rec = 0;
  FILE *probe = std::fopen(path, "rb");
  if (probe) {   // sub_opennext: the file exists -> the record behind the last one
    std::fclose(probe);
    if (!f.read(path, &err)) return rfail(err);
    if (f.dim_len("time") != 0 || f.dim_len("lon") != n_i || f.dim_len("lat") != n_j || f.dim_len("zt") != n_k || !f.var("time") || !f.var("year"))
      return rfail(std::string(path) + ": not a time-slice file of this grid");
    rec = f.numrecs;
  } else {
    // sub_init_netcdf: global attributes, dimensions and axis variables in the reference's order (dimension ids follow it too)
    f.put_att(-1, "Conventions", "CF-1.0");
    f.put_att(-1, "file_name", path);
    f.put_att(-1, "title", "Time averaged integrals");
    if (run_id && *run_id) f.put_att(-1, "experiment_name", run_id);
    f.put_att(-1, "time_unit", "Year mid-point");
    const int d_time = f.add_dim("time", 0), d_xu = f.add_dim("xu", n_i), d_lon = f.add_dim("lon", n_i), d_lat = f.add_dim("lat", n_j),
              d_zt = f.add_dim("zt", n_k), d_yu = f.add_dim("yu", n_j), d_lone = f.add_dim("lon_edges", n_i + 1),
              d_late = f.add_dim("lat_edges", n_j + 1), d_zte = f.add_dim("zt_edges", n_k + 1), d_xue = f.add_dim("xu_edges", n_i + 1),
              d_yue = f.add_dim("yu_edges", n_j + 1);
    f.add_dim("lat_moc", n_j + 1); f.add_dim("zt_moc", n_k + 1); f.add_dim("lat_moc_edges", n_j + 2); f.add_dim("zt_moc_edges", n_k + 2);
    f.add_dim("para", 1);
    defvar(f, "time", cg::NC3_DOUBLE, {d_time}, "T", "Year", "time", "Year mid-point");
    defvar(f, "year", cg::NC3_FLOAT, {d_time}, " ", "year", " ", " ");
    const int v_lon = defvar(f, "lon", cg::NC3_DOUBLE, {d_lon}, "X", "longitude of the t grid", "longitude", "degrees_east");
    const int v_lat = defvar(f, "lat", cg::NC3_DOUBLE, {d_lat}, "Y", "latitude of the t grid", "latitude", "degrees_north");
    const int v_zt = defvar(f, "zt", cg::NC3_DOUBLE, {d_zt}, "Z", "z-level mid depth", "depth", "m");
    const int v_xu = defvar(f, "xu", cg::NC3_DOUBLE, {d_xu}, "X", "longitude of the u grid", "longitude", "degrees_east");
    const int v_yu = defvar(f, "yu", cg::NC3_DOUBLE, {d_yu}, "Y", "latitude of the u grid", "latitude", "degrees_north");
    const int v_lone = defvar(f, "lon_edges", cg::NC3_DOUBLE, {d_lone}, " ", "longitude of t grid edges", " ", "degrees");
    const int v_late = defvar(f, "lat_edges", cg::NC3_DOUBLE, {d_late}, " ", "latitude of t grid edges", " ", "degrees");
    const int v_zte = defvar(f, "zt_edges", cg::NC3_DOUBLE, {d_zte}, " ", "depth of t grid edges", " ", "m");
    const int v_xue = defvar(f, "xu_edges", cg::NC3_DOUBLE, {d_xue}, " ", "longitude of u grid edges", " ", "degrees");
    const int v_yue = defvar(f, "yu_edges", cg::NC3_DOUBLE, {d_yue}, " ", "latitude of u grid edges", " ", "degrees");
    // grid_level 'I' with valid_range (0, 100) as two ints; grid_mask 'F' (0, 100), grid_topo 'F' (0, 5000) as two floats
    const int v_lev = f.add_var("grid_level", cg::NC3_INT, {d_lat, d_lon});
    f.put_att_num(v_lev, "valid_range", cg::NC3_INT, {0.0, 100.0});
    f.put_att_num(v_lev, "missing_value", cg::NC3_DOUBLE, {kNcFillDouble});
    f.put_att(v_lev, "long_name", "grid definition"); f.put_att(v_lev, "standard_name", "model_level_number"); f.put_att(v_lev, "units", "n/a");
    const int v_mask = f.add_var("grid_mask", cg::NC3_FLOAT, {d_lat, d_lon});
    f.put_att_num(v_mask, "valid_range", cg::NC3_FLOAT, {0.0, 100.0});
    f.put_att_num(v_mask, "missing_value", cg::NC3_DOUBLE, {kNcFillDouble});
    f.put_att(v_mask, "long_name", "land-sea mask"); f.put_att(v_mask, "units", "n/a");
    const int v_topo = f.add_var("grid_topo", cg::NC3_FLOAT, {d_lat, d_lon});
    f.put_att_num(v_topo, "valid_range", cg::NC3_FLOAT, {0.0, 5000.0});
    f.put_att_num(v_topo, "missing_value", cg::NC3_DOUBLE, {kNcFillDouble});
    f.put_att(v_topo, "long_name", "ocean depth "); f.put_att(v_topo, "units", "m");
    // sub_save_netcdf, first record only: the axes.  xu = lon_edges(1:n_i), yu = lat_edges(1:n_j); the u-grid edges through
    // edge_maker(2, ...): the t-grid points, then the last point plus the last cell's width (ipo_dlon, ipo_dlat)
    f.put(v_lon, lon, n_i); f.put(v_lone, lon_e, n_i + 1); f.put(v_xu, lon_e, n_i);
    std::vector<double> e(lon, lon + n_i);
    e.push_back(lon[n_i - 1] + 360.0 / n_i);
    f.put(v_xue, e.data(), n_i + 1);
    f.put(v_lat, lat, n_j); f.put(v_late, lat_e, n_j + 1); f.put(v_yu, lat_e, n_j);
    e.assign(lat, lat + n_j);
    e.push_back(lat[n_j - 1] + (lat_e[n_j] - lat_e[n_j - 1]));
    f.put(v_yue, e.data(), n_j + 1);
    f.put(v_zt, zt, n_k); f.put(v_zte, zt_e, n_k + 1);
    std::vector<double> lev((size_t)n_i * n_j), mk((size_t)n_i * n_j), topo((size_t)n_i * n_j);
    for (int j = 0; j < n_j; j++)
      for (int i = 0; i < n_i; i++) {
        const int kk = k1[(i + 1) + (size_t)(n_i + 2) * (j + 1)];
        const bool ocean = kk <= n_k;                       // phys_ocn(ipo_mask_ocn,i,j,n_k) = 1
        lev[i + (size_t)n_i * j] = kk;
        mk[i + (size_t)n_i * j] = ocean ? 1.0 : kNcFillDouble;
        topo[i + (size_t)n_i * j] = ocean ? zt_e[n_k - kk + 1] : kNcFillDouble;   // phys_ocn(ipo_Dbot,i,j,k1(i,j))
      }
    f.put(v_lev, lev.data(), (long long)lev.size()); f.put(v_mask, mk.data(), (long long)mk.size()); f.put(v_topo, topo.data(), (long long)topo.size());
  }
  int d_time = -1, d_lon = -1, d_lat = -1, d_zt = -1;
  for (size_t q = 0; q < f.dims.size(); q++) {
    if (f.dims[q].first == "time") d_time = (int)q;
    if (f.dims[q].first == "lon") d_lon = (int)q;
    if (f.dims[q].first == "lat") d_lat = (int)q;
    if (f.dims[q].first == "zt") d_zt = (int)q;
  }
  const std::vector<int> d4 = {d_time, d_zt, d_lat, d_lon};
  const size_t n3 = (size_t)n_i * n_j * n_k;
  std::vector<double> a(n3);
  auto idv = [&](const char *name) { for (size_t q = 0; q < f.vars.size(); q++) if (f.vars[q].name == name) return (int)q; return -1; };
  const double yr = year_mid, yri = (double)std::llround(year_mid);
  f.put_rec(idv("time"), rec, &yr, 1);
  f.put_rec(idv("year"), rec, &yri, 1);
  // one field: value(l-independent functor) at the wet cells, surface level first, fill value elsewhere (sub_putvar3d_g)
  auto put_field = [&](const std::string &name, const std::string &lname, const std::string &units, double rmin, double rmax, auto &&value) {
    for (int k = 0; k < n_k; k++)
      for (int j = 0; j < n_j; j++)
        for (int i = 0; i < n_i; i++) {
          const int km = n_k - k;
          const bool wet = km >= k1[(i + 1) + (size_t)(n_i + 2) * (j + 1)];
          const size_t cell = (size_t)i + (size_t)n_i * (j + (size_t)n_j * (km - 1));
          a[(size_t)i + (size_t)n_i * (j + (size_t)n_j * k)] = wet ? value(cell) : kNcFillDouble;
        }
    f.put_rec(def_field(f, name, d4, lname, units, rmin, rmax), rec, a.data(), (long long)n3);
  };
  auto ocn = [&](int l, size_t cell) { return int_ocn[l + (size_t)n_ocn * cell]; };
  // ctrl_data_save_slice_ocn, :1981-2022
  int l_dic = -1, l_13 = -1, l_14 = -1, l_s = n_ocn > 1 ? 1 : -1;
  for (int l = 0; l < n_ocn; l++) {
    const std::string nm = ocn_names[l];
    if (nm == "DIC") l_dic = l;
    if (nm == "DIC_13C") l_13 = l;
    if (nm == "DIC_14C") l_14 = l;
    const int ty = ocn_type[l];
    put_field("ocn_" + nm, ocn_longnames[l], ocn_units[l], ocn_mima[2 * l], ocn_mima[2 * l + 1], [&](size_t c) {
      if (ty == 0) return l == 0 ? ocn(l, c) / int_t - 273.15 : ocn(l, c) / int_t;
      if (ty == 1) return ocn(l, c) / int_t;
      const double tot = ocn(ocn_dep[l], c) / int_t, frac = ocn(l, c) / int_t;
      return slice_delta(tot, frac, slice_standard(ty));
    });
  }
  if (l_dic >= 0 && l_13 >= 0 && l_14 >= 0)   // :2023-2046
    put_field("ocn_DIC_D14C", " oceanic D14C (big delta)", "o/oo", 0.0, 0.0, [&](size_t c) {
      const double tot = ocn(l_dic, c) / int_t;
      const double d13 = slice_delta(tot, ocn(l_13, c) / int_t, kStd13C), d14 = slice_delta(tot, ocn(l_14, c) / int_t, kStd14C);
      return 1000.0 * ((1.0 + d14 / 1000.0) * (0.975 * 0.975) / ((1.0 + d13 / 1000.0) * (1.0 + d13 / 1000.0)) - 1.0);
    });
  if (mass && l_s >= 0) {   // ctrl_data_save_derived, :2053-2107
    double sm = 0.0, m = 0.0;   // loc_ocn_mean_S = SUM(int_S * M) / SUM(M) over the whole array (dry cells hold zeros)
    for (size_t c = 0; c < n3; c++) { sm = sm + ocn(l_s, c) * mass[c]; m = m + mass[c]; }
    const double mean_s = sm / m;
    for (int l = 2; l < n_ocn; l++)
      if (ocn_type[l] == 0 || ocn_type[l] == 1)
        put_field(std::string("ocn_") + ocn_names[l] + "_Snorm", std::string(ocn_names[l]) + " normalized by salinity", "mol kg-1", 0.0, 0.0,
                  [&](size_t c) { return ocn(l, c) * (mean_s / ocn(l_s, c)) / int_t; });
    for (int l = 2; l < n_ocn; l++)
      if (ocn_type[l] == 1 || (ocn_type[l] >= 11 && ocn_type[l] <= 22))
        put_field(std::string("ocn_") + ocn_names[l] + "_tot", std::string(ocn_names[l]) + " volume integrated inventory", "mol", 0.0, 0.0,
                  [&](size_t c) { return mass[c] * ocn(l, c) / int_t; });
  }
  for (int ic = 0; ic < n_carb; ic++)   // ctrl_data_save_slice_carb, :2146-2155
    put_field(std::string("carb_") + carb_names[ic], std::string("carbonate chemistry properties - ") + carb_names[ic], " ", 0.0, 0.0,
              [&](size_t c) { return int_carb[ic + (size_t)n_carb * c] / int_t; });
  for (int ic = 0; ic < n_carbconst; ic++)   // ctrl_data_save_slice_carbconst, :2156-2165
    put_field(std::string("carb_const_") + carbconst_names[ic],
              std::string("carbonate chemistry dissociation constants - ") + carbconst_names[ic], " ", 0.0, 0.0,
              [&](size_t c) { return int_carbconst[ic + (size_t)n_carbconst * c] / int_t; });
  if (mass)   // ctrl_data_save_slice_bio .AND. ctrl_data_save_derived, :2167-2199
    for (int l = 0; l < n_sed; l++) {
      const int ty = sed_type[l];
      if (!((ty >= 1 && ty <= 7) || (ty >= 11 && ty <= 22))) continue;   // par_sed_type_age / _frac / _misc have no variable
      put_field(std::string("bio_part_") + sed_names[l], std::string("particulate density - ") + sed_names[l], ty >= 11 ? "o/oo" : "mol kg-1",
                0.0, 0.0, [&](size_t c) {
                  if (ty < 11) return int_part[l + (size_t)n_sed * c] / int_t;
                  return slice_delta(int_part[sed_dep[l] + (size_t)n_sed * c] / int_t, int_part[l + (size_t)n_sed * c] / int_t, slice_standard(ty));
                });
    }
  if (!f.write(path, &err)) return rfail(err);
  return CG_OK;
}


// ------------------------------------------------------------------ ATCHEM restart (ctrl_ncrst = .TRUE., the default)
// sub_data_netCDF_ncrstsave, src/atchem/atchem_data_netCDF.f90:22-109 (title, dimensions lon / lat / lon_edges / lat_edges, one
// FLOAT variable atm_<name>(lat, lon) per selected atmosphere tracer, mask = 1 everywhere: sub_putvar2d, gem_netcdf.f90:664-699).
// atm (n_atm, n_i, n_j) in Fortran order.  lon / lat / lon_e / lat_e: phys_atm's axes through edge_maker
// (atchem_data.f90:195-229); they equal BIOGEM's (biogem_axes in restart.py) on the same grid.
extern "C" int cg_restart_atchem_write(const char *path, int n_i, int n_j, const double *lon, const double *lat, const double *lon_e,
                                       const double *lat_e, int n_atm, const char *const *atm_names,
                                       const char *const *atm_longnames, const double *atm, double year, const char *run_id) {
  if (!path || !lon || !lat || !lon_e || !lat_e || n_i <= 0 || n_j <= 0 || n_atm < 0 || (n_atm && (!atm_names || !atm_longnames || !atm)))
    return rfail("cg_restart_atchem_write: bad argument");
  cg::Nc3File f;
  char y8[16];                                       // CHARACTER(7) receives 8 digits, as in BIOGEM's writer (:34,59)
  std::snprintf(y8, sizeof y8, "%08d", (int)year);
  f.put_att(-1, "Conventions", "CF-1.0");
  f.put_att(-1, "file_name", path);
  f.put_att(-1, "title", std::string("ATCHEM restart @ year ") + std::string(y8).substr(0, 7));
  if (run_id && *run_id) f.put_att(-1, "experiment_name", run_id);
  const int d_lon = f.add_dim("lon", n_i), d_lat = f.add_dim("lat", n_j), d_lone = f.add_dim("lon_edges", n_i + 1),
            d_late = f.add_dim("lat_edges", n_j + 1);
  const int v_lon = defvar(f, "lon", cg::NC3_DOUBLE, {d_lon}, "X", "longitude of the t grid", "longitude", "degrees_east");
  const int v_lat = defvar(f, "lat", cg::NC3_DOUBLE, {d_lat}, "Y", "latitude of the t grid", "latitude", "degrees_north");
  const int v_lone = defvar(f, "lon_edges", cg::NC3_DOUBLE, {d_lone}, " ", "longitude of t grid edges", " ", "degrees");
  const int v_late = defvar(f, "lat_edges", cg::NC3_DOUBLE, {d_late}, " ", "latitude of t grid edges", " ", "degrees");
  std::vector<int> ida(n_atm);
  for (int l = 0; l < n_atm; l++)
    ida[l] = defvar(f, std::string("atm_") + atm_names[l], cg::NC3_FLOAT, {d_lat, d_lon}, " ", atm_longnames[l],
                    std::string("Atmosphere tracer - ") + atm_names[l], " ");
  f.put(v_lon, lon, n_i); f.put(v_lat, lat, n_j); f.put(v_lone, lon_e, n_i + 1); f.put(v_late, lat_e, n_j + 1);
  const size_t n2 = (size_t)n_i * n_j;
  std::vector<double> a(n2);
  for (int l = 0; l < n_atm; l++) {
    for (size_t c = 0; c < n2; c++) a[c] = atm[l + (size_t)n_atm * c];
    f.put(ida[l], a.data(), (long long)n2);
  }
  std::string err;
  if (!f.write(path, &err)) return rfail(err);
  return CG_OK;
}
// sub_data_load_rst, atchem_data.f90:89-189 (netCDF branch) with sub_getvarij (gem_netcdf.f90:1048-1079): every selected tracer
// whose variable atm_<name> is in the file is replaced; the others keep their values.  found (may be NULL): 1 if in the file.
extern "C" int cg_restart_atchem_read(const char *path, int n_i, int n_j, int n_atm, const char *const *atm_names, double *atm,
                                      int32_t *found) {
  if (!path || n_i <= 0 || n_j <= 0 || n_atm < 0 || (n_atm && (!atm_names || !atm))) return rfail("cg_restart_atchem_read: bad argument");
  cg::Nc3File f;
  std::string err;
  if (!f.read(path, &err)) return rfail(err);
  const long long n2 = (long long)n_i * n_j;
  for (int l = 0; l < n_atm; l++) {
    const std::string name = std::string("atm_") + atm_names[l];
    const cg::Nc3Var *v = f.var(name);
    if (found) found[l] = v ? 1 : 0;
    if (!v) continue;
    if (v->count != n2) return rfail(std::string(path) + ": variable " + name + " has the wrong size");
    for (long long c = 0; c < n2; c++) atm[l + (size_t)n_atm * c] = v->data[c];
  }
  return CG_OK;
}

// ------------------------------------------------------------------ binary restarts (ctrl_ncrst = .FALSE.)
// One Fortran unformatted sequential record (gfortran: 4-byte little-endian length before and after the payload; INTEGER is
// 4 bytes, REAL is 8 under -fdefault-real-8, platforms/LINUX:9):
//   ATCHEM  atchem.f90:192-197      n_l_atm, conv_iselected_ia(1:n_l_atm), (atm(ia,:,:), l = 1,n_l_atm)
//   BIOGEM  biogem.f90:2347-2355    n_l_ocn, conv_iselected_io, (ocn(io,:,:,:)), n_l_sed, conv_iselected_is, (bio_part(is,:,:,:))
// ids: the tracers' global indices (ia / io / is of tracer_define.*), which is what the reader matches on
// (atchem_data.f90:176-180, biogem_data.f90:540-547).  Arrays (n_sel, cells) in Fortran order as above, written in full
// (dry cells included, as the reference does).
namespace {
struct FRecord {
  std::vector<unsigned char> b;
  void i32(int32_t v) { unsigned char *p = (unsigned char *)&v; b.insert(b.end(), p, p + 4); }
  void f64(double v) { unsigned char *p = (unsigned char *)&v; b.insert(b.end(), p, p + 8); }
  // tracer l of an (n, cells) array, as the array section a(id,:,:[,:])
  void section(const double *a, int n, int l, size_t cells) { for (size_t c = 0; c < cells; c++) f64(a[l + (size_t)n * c]); }
  bool write(const char *path, std::string *err) const {
    if (b.size() > 0x7fffffffu) { *err = std::string(path) + ": record longer than 2 GiB (gfortran would split it)"; return false; }
    FILE *fp = std::fopen(path, "wb");
    if (!fp) { *err = std::string("cannot open ") + path + " for writing"; return false; }
    const int32_t n = (int32_t)b.size();
    bool ok = std::fwrite(&n, 4, 1, fp) == 1 && (b.empty() || std::fwrite(b.data(), 1, b.size(), fp) == b.size()) && std::fwrite(&n, 4, 1, fp) == 1;
    ok = (std::fclose(fp) == 0) && ok;
    if (!ok) *err = std::string("short write to ") + path;
    return ok;
  }
};
struct FReader {
  std::vector<unsigned char> b;
  size_t pos = 0;
  bool open(const char *path, std::string *err) {
    FILE *fp = std::fopen(path, "rb");
    if (!fp) { *err = std::string("cannot open ") + path; return false; }
    int32_t n = 0, n2 = -1;
    bool ok = std::fread(&n, 4, 1, fp) == 1 && n >= 0;
    if (ok) { b.resize((size_t)n); ok = (n == 0 || std::fread(b.data(), 1, b.size(), fp) == b.size()) && std::fread(&n2, 4, 1, fp) == 1 && n2 == n; }
    std::fclose(fp);
    if (!ok) *err = std::string(path) + ": not a single-record Fortran unformatted file";
    return ok;
  }
  bool i32(int32_t *v) { if (pos + 4 > b.size()) return false; std::memcpy(v, &b[pos], 4); pos += 4; return true; }
  // reads the section of tracer id into column l of an (n, cells) array; l < 0: skipped (a tracer the caller did not select)
  bool section(double *a, int n, int l, size_t cells) {
    if (pos + 8 * cells > b.size()) return false;
    if (l >= 0) for (size_t c = 0; c < cells; c++) std::memcpy(&a[l + (size_t)n * c], &b[pos + 8 * c], 8);
    pos += 8 * cells;
    return true;
  }
};
int find_id(const int32_t *ids, int n, int32_t id) { for (int l = 0; l < n; l++) if (ids[l] == id) return l; return -1; }
// one (count, ids, sections) group of a record into an (n, cells) array
bool read_group(FReader &r, int n, const int32_t *ids, double *a, int32_t *found, size_t cells) {
  int32_t nf = 0;
  if (!r.i32(&nf) || nf < 0 || nf > 4096) return false;
  std::vector<int32_t> fid(nf);
  for (int l = 0; l < nf; l++) if (!r.i32(&fid[l])) return false;
  if (found) for (int l = 0; l < n; l++) found[l] = 0;
  for (int l = 0; l < nf; l++) {
    const int mine = find_id(ids, n, fid[l]);
    if (mine >= 0 && found) found[mine] = 1;
    if (!r.section(a, n, mine, cells)) return false;
  }
  return true;
}
}  // namespace

extern "C" int cg_restart_atchem_write_bin(const char *path, int n_i, int n_j, int n_atm, const int32_t *atm_ids, const double *atm) {
  if (!path || n_i <= 0 || n_j <= 0 || n_atm < 0 || (n_atm && (!atm_ids || !atm))) return rfail("cg_restart_atchem_write_bin: bad argument");
  FRecord r;
  r.i32(n_atm);
  for (int l = 0; l < n_atm; l++) r.i32(atm_ids[l]);
  for (int l = 0; l < n_atm; l++) r.section(atm, n_atm, l, (size_t)n_i * n_j);
  std::string err;
  return r.write(path, &err) ? CG_OK : rfail(err);
}
// The reference reads atm(loc_conv_iselected_ia(l),:,:) for every tracer of the FILE (atchem_data.f90:176-180): tracers of the
// file the caller has not selected are skipped here (the reference stores them in its full-size array, where nothing uses them).
extern "C" int cg_restart_atchem_read_bin(const char *path, int n_i, int n_j, int n_atm, const int32_t *atm_ids, double *atm,
                                          int32_t *found) {
  if (!path || n_i <= 0 || n_j <= 0 || n_atm < 0 || (n_atm && (!atm_ids || !atm))) return rfail("cg_restart_atchem_read_bin: bad argument");
  FReader r;
  std::string err;
  if (!r.open(path, &err)) return rfail(err);
  if (!read_group(r, n_atm, atm_ids, atm, found, (size_t)n_i * n_j) || r.pos != r.b.size())
    return rfail(std::string(path) + ": record does not hold an ATCHEM restart of this grid");
  return CG_OK;
}
extern "C" int cg_restart_biogem_write_bin(const char *path, int n_i, int n_j, int n_k, int n_ocn, const int32_t *ocn_ids,
                                           const double *ocn, int n_sed, const int32_t *sed_ids, const double *bio_part) {
  if (!path || n_i <= 0 || n_j <= 0 || n_k <= 0 || n_ocn < 0 || n_sed < 0 || (n_ocn && (!ocn_ids || !ocn)) || (n_sed && (!sed_ids || !bio_part)))
    return rfail("cg_restart_biogem_write_bin: bad argument");
  const size_t n3 = (size_t)n_i * n_j * n_k;
  FRecord r;
  r.b.reserve(8 + 4 * (size_t)(n_ocn + n_sed) + 8 * n3 * (size_t)(n_ocn + n_sed));
  r.i32(n_ocn);
  for (int l = 0; l < n_ocn; l++) r.i32(ocn_ids[l]);
  for (int l = 0; l < n_ocn; l++) r.section(ocn, n_ocn, l, n3);
  r.i32(n_sed);
  for (int l = 0; l < n_sed; l++) r.i32(sed_ids[l]);
  for (int l = 0; l < n_sed; l++) r.section(bio_part, n_sed, l, n3);
  std::string err;
  return r.write(path, &err) ? CG_OK : rfail(err);
}
extern "C" int cg_restart_biogem_read_bin(const char *path, int n_i, int n_j, int n_k, int n_ocn, const int32_t *ocn_ids, double *ocn,
                                          int32_t *found_ocn, int n_sed, const int32_t *sed_ids, double *bio_part, int32_t *found_sed) {
  if (!path || n_i <= 0 || n_j <= 0 || n_k <= 0 || n_ocn < 0 || n_sed < 0 || (n_ocn && (!ocn_ids || !ocn)) || (n_sed && (!sed_ids || !bio_part)))
    return rfail("cg_restart_biogem_read_bin: bad argument");
  FReader r;
  std::string err;
  if (!r.open(path, &err)) return rfail(err);
  const size_t n3 = (size_t)n_i * n_j * n_k;
  if (!read_group(r, n_ocn, ocn_ids, ocn, found_ocn, n3) || !read_group(r, n_sed, sed_ids, bio_part, found_sed, n3) || r.pos != r.b.size())
    return rfail(std::string(path) + ": record does not hold a BIOGEM restart of this grid");
  return CG_OK;
}

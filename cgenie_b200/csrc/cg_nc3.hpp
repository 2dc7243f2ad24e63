// cg_nc3.hpp -- minimal NetCDF-3 "classic" (CDF-1) codec for the restart files of the physics modules.
//
// The reference writes its restarts through netCDF-Fortran (goldstein_data.f90:153-300, embm_data.f90:83-200,
// gold_seaice_data.f90:100-230); neither that library nor its C core exists in this image, and the files it produces for
// these layouts are plain CDF-1: a header (dimensions, global attributes, variables with their attributes and data
// offsets) followed by the fixed-size variables in definition order, big-endian, each padded to 4 bytes.  This codec
// writes and reads exactly that subset: fixed dimensions, NC_INT / NC_FLOAT / NC_DOUBLE variables, NC_CHAR and numeric
// attributes -- and, for BIOGEM's time-slice files (biogem_data_netCDF.f90:148-459: `time` is unlimited), variables along ONE
// record dimension (a dimension of length 0; the records follow the fixed-size variables, one slab per record variable).  Files written here open with any netCDF library (checked against scipy.io.netcdf_file in
// tests/test_restart_nc.py) and files written by netCDF-3 for these layouts read back here.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace cg {

enum Nc3Type { NC3_CHAR = 2, NC3_INT = 4, NC3_FLOAT = 5, NC3_DOUBLE = 6 };

struct Nc3Att {                                                 // text (type NC3_CHAR) or numeric attribute
  std::string name;
  int type = NC3_CHAR;
  std::string text;
  std::vector<double> num;
};

struct Nc3Var {
  std::string name;
  int type = NC3_DOUBLE;
  std::vector<int> dimids;                                     // file order: slowest first (reverse of the Fortran order)
  std::vector<Nc3Att> atts;                                    // in definition order
  std::vector<double> data;                                    // values, file order (converted on write / read)
  long long count = 0;                                         // product of the dimension lengths (record variables: of one record)
  bool rec = false;                                            // first dimension is the record dimension: data holds numrecs * count values
};

class Nc3File {
 public:
  int add_dim(const std::string &name, int len);
  int add_var(const std::string &name, int type, const std::vector<int> &dimids);   // dimids in FILE order
  void put_att(int varid, const std::string &name, const std::string &value);          // varid -1: global attribute
  void put_att_num(int varid, const std::string &name, int type, const std::vector<double> &v);
  void put(int varid, const double *v, long long n);
  void put(int varid, const int *v, long long n);
  void put_rec(int varid, int rec, const double *v, long long n);                       // record `rec` (0-based) of a record variable; grows numrecs
  bool write(const std::string &path, std::string *err) const;
  bool read(const std::string &path, std::string *err);
  const Nc3Var *var(const std::string &name) const;
  int dim_len(const std::string &name) const;                   // -1 if absent
  std::vector<std::pair<std::string, int>> dims;
  std::vector<Nc3Att> gatts;
  std::vector<Nc3Var> vars;
  int numrecs = 0;
};

}  // namespace cg

// Command-line style options -> Solver_Parameters, accepting the reference's
// sdpb flags for the solver (reference src/sdp_solve/Solver_Parameters/Solver_Parameters.cxx:20-157,
// src/sdpb/SDPB_Parameters.cxx:21-94): "--key=value", "--key value" and bare
// boolean flags.  Unknown options are an error, as with boost::program_options.
#pragma once
#include "solver.hpp"

namespace sdpb_host
{
struct Solve_Options
{
  Solver_Parameters parameters;
  std::string sdp_dir, out_dir;
  int device = 0;
  bool verbose = false;
};
inline Solve_Options parse_options(int argc, const char *const *argv)
{
  Solve_Options o;
  for(int i = 0; i < argc; ++i)
    {
      std::string a = argv[i];
      if(a.rfind("--", 0) != 0)
        throw std::runtime_error("unexpected argument '" + a + "'");
      a = a.substr(2);
      std::string key = a, value;
      bool has_value = false;
      const size_t eq = a.find('=');
      if(eq != std::string::npos)
        {
          key = a.substr(0, eq);
          value = a.substr(eq + 1);
          has_value = true;
        }
      const bool is_flag = key == "findPrimalFeasible" || key == "findDualFeasible"
                           || key == "detectPrimalFeasibleJump" || key == "detectDualFeasibleJump"
                           || key == "noFinalCheckpoint" || key == "verbose";
      if(!has_value && !is_flag)
        {
          if(i + 1 >= argc)
            throw std::runtime_error("option --" + key + " needs a value");
          value = argv[++i];
        }
      if(key == "sdpDir" || key == "s")
        o.sdp_dir = value;
      else if(key == "outDir" || key == "o")
        o.out_dir = value;
      else if(key == "device")
        o.device = std::stoi(value);
      else if(key == "verbose")
        o.verbose = true;
      else if(!o.parameters.set(key, value))
        throw std::runtime_error("unrecognised option '--" + key + "'");
    }
  if(o.sdp_dir.empty())
    throw std::runtime_error("the option '--sdpDir' is required but missing");
  return o;
}
} // namespace sdpb_host

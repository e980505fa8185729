// `sdpb_b200_solve` executable: the reference's `sdpb` command line for the
// solver options (reference src/sdpb/main.cxx:31-190), single process, one GPU.
#include "../../../include/sdpb_b200_solver.h"

#include <cstdio>

int main(int argc, char **argv)
{
  char summary[4096];
  const int rc = sdpb_b200_solve(argc - 1, argv + 1, summary, sizeof summary);
  fprintf(rc ? stderr : stdout, "%s\n", summary);
  return rc;
}

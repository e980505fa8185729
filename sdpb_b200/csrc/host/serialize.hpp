// El::BigFloat::Serialize / Deserialize and a few file-system helpers, shared by the checkpoint
// code (checkpoint.hpp) and the binary SDP reader (block_data_bin.hpp).
//
// El::BigFloat::Serialize belongs to the un-vendored Elemental fork (Dockerfile:30).  The image
// restated here is the one that fork's GMP-backed BigFloat writes: the three scalar fields of the
// __mpf_struct followed by the limb array at its allocated length,
//     int32 _mp_prec | int32 _mp_size | int64 _mp_exp | (_mp_prec + 1) limbs, little-endian,
// 16 + 8 (prec_limbs + 1) bytes -- 128 at --precision=768 (sdpb_util/memory_estimates.cxx:9-12
// counts the same limbs).  No serialized BigFloat ships with the reference: this layout is NOT
// pinned by a fixture.
#pragma once
#include "bigfloat.hpp"

#include <cstdio>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

namespace sdpb_host
{
inline size_t bigfloat_serialized_size() { return 16 + 8 * (size_t)(prec_limbs() + 1); }
inline void serialize_bigfloat(const BigFloat &f, uint8_t *buf)
{
  const int32_t prec = (int32_t)f.v[0]._mp_prec, size = (int32_t)f.v[0]._mp_size;
  const int64_t exp = (int64_t)f.v[0]._mp_exp;
  memcpy(buf, &prec, 4);
  memcpy(buf + 4, &size, 4);
  memcpy(buf + 8, &exp, 8);
  const int alloc = prec_limbs() + 1, used = size < 0 ? -size : size;
  for(int i = 0; i < alloc; ++i)
    {
      const uint64_t limb = i < used ? (uint64_t)f.v[0]._mp_d[i] : 0; // limbs beyond _mp_size are garbage in GMP
      memcpy(buf + 16 + 8 * (size_t)i, &limb, 8);
    }
}
inline void deserialize_bigfloat(BigFloat &f, const uint8_t *buf)
{
  int32_t prec, size;
  int64_t exp;
  memcpy(&prec, buf, 4);
  memcpy(&size, buf + 4, 4);
  memcpy(&exp, buf + 8, 8);
  const int used = size < 0 ? -size : size;
  if(prec != (int32_t)f.v[0]._mp_prec || used > prec + 1)
    throw std::runtime_error("Corrupted binary checkpoint file: element written at another precision ("
                             + std::to_string(prec) + " limbs, expected " + std::to_string((int)f.v[0]._mp_prec) + ")");
  for(int i = 0; i < used; ++i)
    {
      uint64_t limb;
      memcpy(&limb, buf + 16 + 8 * (size_t)i, 8);
      f.v[0]._mp_d[i] = limb;
    }
  f.v[0]._mp_size = size;
  f.v[0]._mp_exp = (long)exp;
}

inline bool path_exists(const std::string &p)
{
  struct stat st;
  return stat(p.c_str(), &st) == 0;
}
inline bool is_directory(const std::string &p)
{
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
inline void create_directories(const std::string &p)
{
  for(size_t i = 1; i <= p.size(); ++i)
    if(i == p.size() || p[i] == '/')
      mkdir(p.substr(0, i).c_str(), 0777);
}

} // namespace sdpb_host

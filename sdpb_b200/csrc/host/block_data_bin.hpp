// The binary form of an SDP block, block_data_<j>.bin: what stock pmp2sdp
// writes by default and sdpb reads at
//   parse_block_data                  src/sdp_solve/SDP/read_block_data/SDP_Block_Data.cxx:32-48
//   serialize(El::BigFloat), save/load(El::Matrix)   src/sdpb_util/boost_serialization.hxx:17-97
//   (writer: src/pmp2sdp/write_block_data.cxx, "NB: this should match")
// i.e. a boost::archive::binary_oarchive holding, in this order,
//     mp_bitcnt_t precision | El::Matrix B | std::vector<BigFloat> c | Matrix bases_even | Matrix bases_odd.
//
// Neither Boost nor Elemental is in this image, and the reference ships no .bin sample, so the
// byte layout is restated from Boost.Serialization's binary archive (library version >= 8, i.e.
// Boost >= 1.44 -- the sizes basic_binary_iarchive.hpp branches on) and from the serialize()
// bodies cited above; it is NOT pinned by a fixture.  Layout:
//     u64 22, "serialization::archive", u16 library_version                     archive header
//     u64 precision                                                             primitive
//     Matrix:  [u8 tracking = 0, u32 class version = 0 -- first Matrix of the archive only]
//              Int height, Int width, Int ldim (El::Int: 4 bytes, or 8 with EL_USE_64BIT_INTS;
//              both are accepted, told apart by the shape block_info promises),
//              ldim * width BigFloat items, column-major
//     BigFloat:[u8 tracking = 0, u32 class version = 1 -- first BigFloat of the archive only]
//              u8 is_zero; unless zero: the El::BigFloat::Serialize image (checkpoint.hpp),
//              as a raw byte array without a length
//     vector:  u64 count, u32 item_version, count BigFloat items (std::vector is
//              object_serializable: no class preamble)
// The writer below produces exactly this; tests convert the reference's JSON fixtures, delete the
// JSON and replay the golden trajectories from the .bin files.
#pragma once
#include "serialize.hpp"

namespace sdpb_host
{
class Boost_Binary_Reader
{
  std::ifstream in;
  std::string path;
  bool seen_bigfloat = false, seen_matrix = false;
  std::vector<uint8_t> buf;
  void raw(void *p, size_t n)
  {
    in.read(reinterpret_cast<char *>(p), (std::streamsize)n);
    if(!in.good())
      throw std::runtime_error("Unexpected end of binary block data: " + path);
  }

public:
  int int_bytes = 4; // sizeof(El::Int) of the writer
  unsigned library_version = 0;
  explicit Boost_Binary_Reader(const std::string &p) : in(p, std::ios::binary), path(p), buf(bigfloat_serialized_size())
  {
    if(!in)
      throw std::runtime_error("Unable to open: " + path);
    const uint64_t len = u64();
    char sig[32] = {0};
    if(len != 22)
      throw std::runtime_error("Not a Boost binary archive: " + path);
    raw(sig, 22);
    if(std::string(sig, 22) != "serialization::archive")
      throw std::runtime_error("Not a Boost binary archive: " + path);
    uint16_t v;
    raw(&v, 2);
    library_version = v;
    if(library_version < 8)
      throw std::runtime_error("Boost binary archive of library version " + std::to_string(v)
                               + " (< 8) is not supported: " + path);
  }
  uint64_t u64()
  {
    uint64_t v;
    raw(&v, 8);
    return v;
  }
  uint32_t u32()
  {
    uint32_t v;
    raw(&v, 4);
    return v;
  }
  uint8_t u8()
  {
    uint8_t v;
    raw(&v, 1);
    return v;
  }
  std::streampos tell() { return in.tellg(); }
  void seek(std::streampos p) { in.seekg(p); }
  void bigfloat(BigFloat &f)
  {
    if(!seen_bigfloat)
      {
        const uint8_t tracking = u8();
        const uint32_t version = u32();
        if(tracking != 0 || version != 1)
          throw std::runtime_error("Unexpected class preamble of El::BigFloat (tracking " + std::to_string(tracking)
                                   + ", version " + std::to_string(version) + "): " + path);
        seen_bigfloat = true;
      }
    if(u8()) // is_zero (boost_serialization.hxx:29-42)
      {
        f.zero();
        return;
      }
    raw(buf.data(), buf.size());
    deserialize_bigfloat(f, buf.data());
  }
  // expect_h / expect_w < 0: unknown
  void matrix(Matrix &m, long expect_h, long expect_w)
  {
    const bool first = !seen_matrix;
    if(first)
      {
        const uint8_t tracking = u8();
        const uint32_t version = u32();
        if(tracking != 0 || version != 0)
          throw std::runtime_error("Unexpected class preamble of El::Matrix: " + path);
        seen_matrix = true;
      }
    const std::streampos at = tell();
    long h = 0, w = 0, ld = 0;
    auto read_dims = [&]() {
      if(int_bytes == 4)
        {
          h = (int32_t)u32();
          w = (int32_t)u32();
          ld = (int32_t)u32();
        }
      else
        {
          h = (long)u64();
          w = (long)u64();
          ld = (long)u64();
        }
    };
    auto plausible = [&]() {
      return h >= 0 && w >= 0 && ld >= h && ld >= 1 && (expect_h < 0 || h == expect_h) && (expect_w < 0 || w == expect_w);
    };
    read_dims();
    if(!plausible() && first)
      {
        int_bytes = 8; // a build of Elemental with 64-bit El::Int
        seek(at);
        read_dims();
      }
    if(!plausible())
      throw std::runtime_error("block data matrix has the wrong size: (" + std::to_string(h) + "," + std::to_string(w)
                               + "), leading dimension " + std::to_string(ld) + ", expected (" + std::to_string(expect_h)
                               + "," + std::to_string(expect_w) + "): " + path);
    m.resize((int)h, (int)w);
    BigFloat pad;
    for(long j = 0; j < w; ++j)
      for(long i = 0; i < ld; ++i)
        bigfloat(i < h ? m((int)i, (int)j) : pad);
  }
  void vector(Matrix &v)
  {
    const uint64_t count = u64();
    u32(); // item_version
    if(count > (uint64_t)1 << 40)
      throw std::runtime_error("Corrupted vector length in " + path);
    v.resize((int)count, 1);
    for(uint64_t i = 0; i < count; ++i)
      bigfloat(v((int)i, 0));
  }
};

class Boost_Binary_Writer
{
  std::ofstream out;
  bool seen_bigfloat = false, seen_matrix = false;
  std::vector<uint8_t> buf;
  void raw(const void *p, size_t n) { out.write(reinterpret_cast<const char *>(p), (std::streamsize)n); }

public:
  explicit Boost_Binary_Writer(const std::string &path, uint16_t library_version = 19)
      : out(path, std::ios::binary), buf(bigfloat_serialized_size())
  {
    const uint64_t len = 22;
    raw(&len, 8);
    raw("serialization::archive", 22);
    raw(&library_version, 2);
  }
  void u64(uint64_t v) { raw(&v, 8); }
  void bigfloat(const BigFloat &f)
  {
    if(!seen_bigfloat)
      {
        const uint8_t tracking = 0;
        const uint32_t version = 1;
        raw(&tracking, 1);
        raw(&version, 4);
        seen_bigfloat = true;
      }
    const uint8_t is_zero = f.sgn() == 0;
    raw(&is_zero, 1);
    if(is_zero)
      return;
    serialize_bigfloat(f, buf.data());
    raw(buf.data(), buf.size());
  }
  void matrix(const Matrix &m)
  {
    if(!seen_matrix)
      {
        const uint8_t tracking = 0;
        const uint32_t version = 0;
        raw(&tracking, 1);
        raw(&version, 4);
        seen_matrix = true;
      }
    const int32_t h = m.h, w = m.w, ld = m.h > 1 ? m.h : 1; // El::Matrix: LDim = max(height, 1)
    raw(&h, 4);
    raw(&w, 4);
    raw(&ld, 4);
    const BigFloat pad;
    for(int j = 0; j < w; ++j)
      for(int i = 0; i < ld; ++i)
        bigfloat(i < h ? m(i, j) : pad);
  }
  void vector(const Matrix &v)
  {
    u64((uint64_t)v.h);
    const uint32_t item_version = 1;
    raw(&item_version, 4);
    for(int i = 0; i < v.h; ++i)
      bigfloat(v(i, 0));
  }
  bool good() { return out.good(); }
};

// parse_block_data, Block_File_Format::bin (SDP_Block_Data.cxx:36-48).  P, N, n, h_even, h_odd: the
// shapes block_info_<j>.json and objectives.json promise.
inline void read_block_data_bin(const std::string &path, long P, long N, long n, long h_even, long h_odd, Matrix &B,
                                Matrix &c, Matrix &bases_even, Matrix &bases_odd)
{
  Boost_Binary_Reader ar(path);
  const uint64_t precision = ar.u64();
  if(precision != (uint64_t)working_precision_bits())
    throw std::runtime_error("Read GMP precision: " + std::to_string(precision)
                             + ", expected: " + std::to_string(working_precision_bits()) + " in " + path);
  ar.matrix(B, P, N);
  ar.vector(c);
  ar.matrix(bases_even, h_even, n);
  ar.matrix(bases_odd, h_odd, h_odd ? n : -1);
}
inline void write_block_data_bin(const std::string &path, const Matrix &B, const Matrix &c, const Matrix &bases_even,
                                 const Matrix &bases_odd)
{
  Boost_Binary_Writer ar(path);
  ar.u64((uint64_t)working_precision_bits());
  ar.matrix(B);
  ar.vector(c);
  ar.matrix(bases_even);
  ar.matrix(bases_odd);
  if(!ar.good())
    throw std::runtime_error("Error when writing to: " + path);
}
} // namespace sdpb_host

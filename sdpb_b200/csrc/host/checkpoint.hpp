// Checkpoints of the solver state x, X, y, Y:
//   SDP_Solver::save_checkpoint      src/sdp_solve/SDP_Solver/save_checkpoint.cxx:13-123
//   load_binary_checkpoint           src/sdp_solve/SDP_Solver/load_checkpoint/load_binary_checkpoint.cxx:9-152
//   load_text_checkpoint             src/sdp_solve/SDP_Solver/load_checkpoint/load_text_checkpoint.cxx:6-46
//   read_text_block                  src/sdp_solve/read_text_block.hxx:24-58
//   SDP_Solver::load_checkpoint      src/sdp_solve/SDP_Solver/load_checkpoint/load_checkpoint.cxx:14-30
//
// Binary layout (save_checkpoint.cxx:13-36), one file "checkpoint_<generation>_<rank>" per rank:
// for each of x, X, y, Y and each of its local blocks
//     int64 local_height, int64 local_width,
//     local_height * local_width elements, ROW by row (the loops of :26-27), each an
//     El::BigFloat::Serialize image of SerializedSize() bytes;
// next to it "checkpoint.json" = {"current": g, "backup": g', "version": ..., "options": {...}}
// (:102-122), written as checkpoint_new.json and renamed.  y is stored once per block (the
// reference keeps one copy of y per block, SDP_Solver.cxx:12-14); all copies are equal.
//
// El::BigFloat::Serialize belongs to the un-vendored Elemental fork (Dockerfile:30).  The image
// restated here is the one that fork's GMP-backed BigFloat writes: the three scalar fields of the
// __mpf_struct followed by the limb array at its allocated length,
//     int32 _mp_prec | int32 _mp_size | int64 _mp_exp | (_mp_prec + 1) limbs, little-endian,
// 16 + 8 (prec_limbs + 1) bytes -- 128 at --precision=768 (sdpb_util/memory_estimates.cxx:9-12
// counts the same limbs).  No binary checkpoint ships with the reference, so this layout is NOT
// pinned by a fixture: what is tested is the round trip (a run restarted from its checkpoint
// continues bit for bit) and the text route, whose format the reference tree does define.
#pragma once
#include "serialize.hpp"

namespace sdpb_host
{
// the solver state a checkpoint holds; y_copies: how many block copies of y the reference writes
struct Checkpoint_State
{
  std::vector<Matrix> *x, *X, *Y;
  Matrix *y;
  long current_generation = 0;
  long backup_generation = -1; // -1: none
};

inline void write_block(const Matrix &block, std::ofstream &out, std::vector<uint8_t> &buf)
{
  const int64_t h = block.h, w = block.w;
  out.write(reinterpret_cast<const char *>(&h), 8);
  out.write(reinterpret_cast<const char *>(&w), 8);
  for(int64_t row = 0; row < h; ++row)
    for(int64_t col = 0; col < w; ++col)
      {
        serialize_bigfloat(block((int)row, (int)col), buf.data());
        out.write(reinterpret_cast<const char *>(buf.data()), (std::streamsize)buf.size());
      }
}
// save_checkpoint.cxx:38-123 (single rank: file suffix _0).  options_json: the body of "options".
inline void save_checkpoint(const std::string &dir, Checkpoint_State &st, const std::string &options_json,
                            const std::string &version)
{
  if(dir.empty())
    return;
  if(!path_exists(dir))
    create_directories(dir);
  else if(!is_directory(dir))
    throw std::runtime_error("Checkpoint directory already exists, but is not a directory: " + dir);
  if(st.backup_generation >= 0)
    remove((dir + "/checkpoint_" + std::to_string(st.backup_generation) + "_0").c_str());
  st.backup_generation = st.current_generation;
  st.current_generation += 1;
  const std::string name = dir + "/checkpoint_" + std::to_string(st.current_generation) + "_0";
  std::vector<uint8_t> buf(bigfloat_serialized_size());
  bool wrote = false;
  for(int attempt = 0; attempt < 10 && !wrote; ++attempt)
    {
      std::ofstream out(name, std::ios::binary);
      for(const Matrix &b : *st.x)
        write_block(b, out, buf);
      for(const Matrix &b : *st.X)
        write_block(b, out, buf);
      for(size_t j = 0; j < st.x->size(); ++j)
        write_block(*st.y, out, buf);
      for(const Matrix &b : *st.Y)
        write_block(b, out, buf);
      wrote = out.good();
      if(!wrote && attempt == 9)
        throw std::runtime_error("Error writing checkpoint file " + name + ":  Exceeded max retries.");
    }
  {
    std::ofstream meta(dir + "/checkpoint_new.json");
    meta << "{\n    \"current\": " << st.current_generation << ",\n    \"backup\": " << st.backup_generation
         << ",\n    \"version\": \"" << version << "\",\n    \"options\": \n" << options_json << "}\n";
  }
  rename((dir + "/checkpoint_new.json").c_str(), (dir + "/checkpoint.json").c_str());
}

inline void read_block(Matrix &block, std::ifstream &in, std::vector<uint8_t> &buf)
{
  int64_t h = 0, w = 0;
  in.read(reinterpret_cast<char *>(&h), 8);
  in.read(reinterpret_cast<char *>(&w), 8);
  const std::string dims = "(" + std::to_string(block.h) + "," + std::to_string(block.w) + ")";
  if(!in.good())
    throw std::runtime_error("Corrupted binary checkpoint file.  For block with global size " + dims
                             + ", error when reading height and width");
  if(h != block.h || w != block.w)
    throw std::runtime_error("Incompatible binary checkpoint file.  For block with global size " + dims
                             + ", expected local dimensions " + dims + ", but found (" + std::to_string(h) + ","
                             + std::to_string(w) + ")");
  for(int64_t row = 0; row < h; ++row)
    for(int64_t col = 0; col < w; ++col)
      {
        in.read(reinterpret_cast<char *>(buf.data()), (std::streamsize)buf.size());
        if(!in.good())
          throw std::runtime_error("Corrupted binary checkpoint file. For block with global size " + dims
                                   + ", error when reading element (" + std::to_string(row) + ","
                                   + std::to_string(col) + ")");
        deserialize_bigfloat(block((int)row, (int)col), buf.data());
      }
}
// "key": <integer> of the flat metadata object (boost::property_tree reads it the same way)
inline bool json_integer(const std::string &text, const std::string &key, long &value)
{
  const size_t k = text.find("\"" + key + "\"");
  if(k == std::string::npos)
    return false;
  size_t p = text.find(':', k);
  if(p == std::string::npos)
    return false;
  ++p;
  while(p < text.size() && (isspace((unsigned char)text[p]) || text[p] == '"'))
    ++p;
  char *end = nullptr;
  value = strtol(text.c_str() + p, &end, 10);
  return end != text.c_str() + p;
}
// load_binary_checkpoint.cxx:59-152
inline bool load_binary_checkpoint(const std::string &dir, Checkpoint_State &st)
{
  long current = -1, backup = -1;
  const std::string metadata = dir + "/checkpoint.json";
  if(path_exists(metadata))
    {
      std::ifstream f(metadata);
      std::stringstream ss;
      ss << f.rdbuf();
      if(!json_integer(ss.str(), "current", current))
        throw std::runtime_error("Invalid or missing element 'current' in " + metadata);
      json_integer(ss.str(), "backup", backup);
    }
  std::string name;
  if(current != -1)
    {
      name = dir + "/checkpoint_" + std::to_string(current) + "_0";
      if(!path_exists(name))
        throw std::runtime_error("Missing checkpoint file: " + name);
    }
  else
    {
      name = dir + "/checkpoint.0"; // the pre-generation naming (:121-123)
      if(!path_exists(name))
        return false;
      current = 0;
    }
  std::ifstream in(name, std::ios::binary);
  std::vector<uint8_t> buf(bigfloat_serialized_size());
  for(Matrix &b : *st.x)
    read_block(b, in, buf);
  for(Matrix &b : *st.X)
    read_block(b, in, buf);
  for(size_t j = 0; j < st.x->size(); ++j)
    read_block(*st.y, in, buf);
  for(Matrix &b : *st.Y)
    read_block(b, in, buf);
  st.current_generation = current;
  st.backup_generation = backup;
  return true;
}

// read_text_block.hxx:24-58: "height width" then the elements row by row, one decimal per token
inline void read_text_block(Matrix &block, const std::string &path)
{
  std::ifstream in(path);
  if(!in)
    throw std::runtime_error("Unable to open checkpoint file: " + path);
  long h, w;
  in >> h >> w;
  if(!in.good())
    throw std::runtime_error("Corrupted header in file: " + path);
  if(h != block.h || w != block.w)
    throw std::runtime_error("Incompatible checkpoint file: " + path + ":  Expected dimensions ("
                             + std::to_string(block.h) + "," + std::to_string(block.w) + "), but found ("
                             + std::to_string(h) + "," + std::to_string(w) + ")");
  std::string element;
  for(long row = 0; row < h; ++row)
    for(long col = 0; col < w; ++col)
      {
        in >> element;
        if(in.fail())
          throw std::runtime_error("Corrupted data in file: " + path);
        block((int)row, (int)col) = BigFloat(element);
      }
}
// load_text_checkpoint.cxx:6-46: the files --writeSolution=x,y,X,Y leaves in an out directory
inline bool load_text_checkpoint(const std::string &dir, Checkpoint_State &st)
{
  if(!path_exists(dir + "/x_0.txt"))
    return false;
  for(size_t j = 0; j < st.x->size(); ++j)
    {
      read_text_block((*st.x)[j], dir + "/x_" + std::to_string(j) + ".txt");
      for(int parity = 0; parity < 2; ++parity)
        if((*st.X)[2 * j + parity].h != 0) // constant constraints have empty odd-parity blocks
          {
            const std::string suffix = std::to_string(2 * j + parity) + ".txt";
            read_text_block((*st.X)[2 * j + parity], dir + "/X_matrix_" + suffix);
            read_text_block((*st.Y)[2 * j + parity], dir + "/Y_matrix_" + suffix);
          }
    }
  read_text_block(*st.y, dir + "/y.txt");
  return true;
}
// load_checkpoint.cxx:14-30
inline bool load_checkpoint(const std::string &dir, Checkpoint_State &st, bool require_initial_checkpoint)
{
  const bool valid = !dir.empty() && is_directory(dir) && (load_binary_checkpoint(dir, st) || load_text_checkpoint(dir, st));
  if(!valid && require_initial_checkpoint)
    throw std::runtime_error("Unable to load checkpoint from directory: " + dir);
  return valid;
}
} // namespace sdpb_host

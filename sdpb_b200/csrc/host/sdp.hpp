// Host-side SDP data model and reader: the JSON form of SDPB's `sdp/`
// directory (reference docs/SDPB_input_format.md:16-35; readers at
// src/sdp_solve/SDP/read_objectives.cxx, read_block_data/Json_Block_Data_Parser.hxx:26-36,
// src/sdp_solve/Block_Info/read_block_info.cxx:15-40).  Decimal strings are
// parsed with mpf_set_str at the working precision, exactly what
// El::BigFloat(string) does in the reference.  Block data may also come in the
// binary (Boost serialization) form, block_data_<j>.bin (block_data_bin.hpp);
// an SDP packed by `pmp2sdp --zip` (a stored zip, zip_store.hpp) is unpacked first.
#pragma once
#include "bigfloat.hpp"
#include "block_data_bin.hpp"
#include "zip_store.hpp"

#include <cctype>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace sdpb_host
{
// ---------------------------------------------------------------- tiny JSON
struct Json
{
  enum Kind
  {
    Null,
    Bool,
    Number,
    String,
    Array,
    Object
  } kind
    = Null;
  std::string text; // String / Number (as written) / Bool ("true"/"false")
  std::vector<Json> items;
  std::vector<std::pair<std::string, Json>> members;
  const Json *find(const std::string &key) const
  {
    for(const auto &m : members)
      if(m.first == key)
        return &m.second;
    return nullptr;
  }
  const Json &at(const std::string &key) const
  {
    const Json *p = find(key);
    if(!p)
      throw std::runtime_error("JSON: missing key '" + key + "'");
    return *p;
  }
};

class Json_Parser
{
  const std::string &s;
  size_t i = 0;
  void ws()
  {
    while(i < s.size() && std::isspace((unsigned char)s[i]))
      ++i;
  }
  [[noreturn]] void fail(const std::string &what)
  {
    throw std::runtime_error("JSON parse error at byte " + std::to_string(i) + ": " + what);
  }
  std::string string()
  {
    std::string out;
    ++i; // opening quote
    while(i < s.size() && s[i] != '"')
      {
        if(s[i] == '\\' && i + 1 < s.size())
          {
            ++i;
            switch(s[i])
              {
              case 'n': out += '\n'; break;
              case 't': out += '\t'; break;
              case 'r': out += '\r'; break;
              case 'b': out += '\b'; break;
              case 'f': out += '\f'; break;
              case 'u':
                out += '?';
                i += 4;
                break;
              default: out += s[i];
              }
            ++i;
          }
        else
          out += s[i++];
      }
    if(i >= s.size())
      fail("unterminated string");
    ++i;
    return out;
  }

public:
  explicit Json_Parser(const std::string &text) : s(text) {}
  Json value()
  {
    ws();
    if(i >= s.size())
      fail("unexpected end");
    Json v;
    const char c = s[i];
    if(c == '{')
      {
        v.kind = Json::Object;
        ++i;
        ws();
        if(i < s.size() && s[i] == '}')
          {
            ++i;
            return v;
          }
        for(;;)
          {
            ws();
            if(i >= s.size() || s[i] != '"')
              fail("expected a key");
            std::string key = string();
            ws();
            if(i >= s.size() || s[i] != ':')
              fail("expected ':'");
            ++i;
            v.members.emplace_back(std::move(key), value());
            ws();
            if(i < s.size() && s[i] == ',')
              {
                ++i;
                continue;
              }
            if(i < s.size() && s[i] == '}')
              {
                ++i;
                return v;
              }
            fail("expected ',' or '}'");
          }
      }
    if(c == '[')
      {
        v.kind = Json::Array;
        ++i;
        ws();
        if(i < s.size() && s[i] == ']')
          {
            ++i;
            return v;
          }
        for(;;)
          {
            v.items.push_back(value());
            ws();
            if(i < s.size() && s[i] == ',')
              {
                ++i;
                continue;
              }
            if(i < s.size() && s[i] == ']')
              {
                ++i;
                return v;
              }
            fail("expected ',' or ']'");
          }
      }
    if(c == '"')
      {
        v.kind = Json::String;
        v.text = string();
        return v;
      }
    const size_t b = i;
    while(i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '-' || s[i] == '+' || s[i] == '.'))
      ++i;
    if(i == b)
      fail("unexpected character");
    v.text = s.substr(b, i - b);
    if(v.text == "null")
      v.kind = Json::Null;
    else if(v.text == "true" || v.text == "false")
      v.kind = Json::Bool;
    else
      v.kind = Json::Number;
    return v;
  }
};

inline std::string read_file(const std::string &path)
{
  std::ifstream f(path, std::ios::binary);
  if(!f.good())
    throw std::runtime_error("Cannot read '" + path + "'");
  std::ostringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
inline bool file_exists(const std::string &path)
{
  std::ifstream f(path);
  return f.good();
}
inline Json read_json(const std::string &path)
{
  const std::string text = read_file(path);
  try
    {
      return Json_Parser(text).value();
    }
  catch(std::exception &e)
    {
      throw std::runtime_error(path + ": " + e.what());
    }
}

// ------------------------------------------------------------------ shapes
// Block_Info (reference src/sdp_solve/Block_Info.hxx:14-131), single process:
// every block is local, block_indices = 0..J-1.
struct Block_Info
{
  std::vector<int> dimensions; // m_j   (Block_Info.hxx:23)
  std::vector<int> num_points; // n_j = degree + 1 (Block_Info.hxx:24)
  int num_blocks() const { return (int)dimensions.size(); }
  int schur_block_size(int j) const // Block_Info.hxx:54-58
  {
    return num_points[j] * dimensions[j] * (dimensions[j] + 1) / 2;
  }
  int bilinear_pairing_block_size(int j) const { return num_points[j] * dimensions[j]; } // :69-74
  int psd_matrix_block_size(int j, int parity) const // :83-95
  {
    const int even = dimensions[j] * ((num_points[j] + 1) / 2);
    return parity == 0 ? even : dimensions[j] * num_points[j] - even;
  }
  int bilinear_bases_height(int j, int parity) const // :110-115
  {
    const int degree = num_points[j] - 1;
    return (degree + parity) / 2 + 1 - parity;
  }
  size_t total_psd_rows() const // run.cxx:243-245
  {
    size_t t = 0;
    for(int j = 0; j < num_blocks(); ++j)
      t += (size_t)psd_matrix_block_size(j, 0) + psd_matrix_block_size(j, 1);
    return t;
  }
};

// SDP (reference src/sdp_solve/SDP.hxx:74-122)
struct SDP
{
  BigFloat objective_const;
  Matrix dual_objective_b;                // N x 1
  std::vector<Matrix> primal_objective_c; // J, P_j x 1
  std::vector<Matrix> free_var_matrix;    // J, P_j x N   (B)
  std::vector<Matrix> bilinear_bases;     // 2J, h_p x n_j
  std::vector<BigFloat> normalization;    // optional (N+1), used only for z.txt
  int N() const { return dual_objective_b.h; }
};

inline BigFloat json_number(const Json &v)
{
  if(v.kind != Json::String && v.kind != Json::Number)
    throw std::runtime_error("JSON: expected a number");
  return BigFloat(v.text);
}
inline void json_vector(const Json &v, Matrix &out)
{
  out.resize((int)v.items.size(), 1);
  for(size_t i = 0; i < v.items.size(); ++i)
    out((int)i, 0) = json_number(v.items[i]);
}
// array of rows
inline void json_matrix(const Json &v, Matrix &out, int width_if_empty)
{
  const int h = (int)v.items.size();
  const int w = h ? (int)v.items[0].items.size() : width_if_empty;
  out.resize(h, w);
  for(int i = 0; i < h; ++i)
    {
      if((int)v.items[i].items.size() != w)
        throw std::runtime_error("JSON: ragged matrix");
      for(int j = 0; j < w; ++j)
        out(i, j) = json_number(v.items[i].items[j]);
    }
}

// Read `sdp_dir`: a plain directory, or the zip archive `pmp2sdp --zip` packs it into.
inline void read_sdp(const std::string &sdp_path, Block_Info &block_info, SDP &sdp)
{
  std::unique_ptr<Extracted_Zip> unpacked;
  if(is_regular_file(sdp_path))
    unpacked.reset(new Extracted_Zip(sdp_path));
  const std::string sdp_dir = unpacked ? unpacked->dir : sdp_path;
  const Json control = read_json(sdp_dir + "/control.json");
  const int J = std::stoi(control.at("num_blocks").text);
  {
    const Json obj = read_json(sdp_dir + "/objectives.json");
    sdp.objective_const = json_number(obj.at("constant"));
    json_vector(obj.at("b"), sdp.dual_objective_b);
  }
  sdp.normalization.clear();
  if(file_exists(sdp_dir + "/normalization.json"))
    {
      const Json nj = read_json(sdp_dir + "/normalization.json");
      for(const auto &e : nj.at("normalization").items)
        sdp.normalization.push_back(json_number(e));
    }
  block_info.dimensions.assign(J, 0);
  block_info.num_points.assign(J, 0);
  sdp.primal_objective_c.assign(J, Matrix());
  sdp.free_var_matrix.assign(J, Matrix());
  sdp.bilinear_bases.assign(2 * J, Matrix());
  for(int j = 0; j < J; ++j)
    {
      const std::string suffix = "_" + std::to_string(j) + ".json";
      const std::string bin = sdp_dir + "/block_data_" + std::to_string(j) + ".bin";
      const Json info = read_json(sdp_dir + "/block_info" + suffix);
      block_info.dimensions[j] = std::stoi(info.at("dim").text);
      block_info.num_points[j] = std::stoi(info.at("num_points").text);
      const int n = block_info.num_points[j];
      // read_block_data.cxx: block_data_<j>.bin (what pmp2sdp writes by default) or .json
      if(file_exists(bin))
        read_block_data_bin(bin, block_info.schur_block_size(j), sdp.N(), n, block_info.bilinear_bases_height(j, 0),
                            block_info.bilinear_bases_height(j, 1), sdp.free_var_matrix[j], sdp.primal_objective_c[j],
                            sdp.bilinear_bases[2 * j], sdp.bilinear_bases[2 * j + 1]);
      else if(file_exists(sdp_dir + "/block_data" + suffix))
        {
          const Json data = read_json(sdp_dir + "/block_data" + suffix);
          json_matrix(data.at("bilinear_bases_even"), sdp.bilinear_bases[2 * j], n);
          json_matrix(data.at("bilinear_bases_odd"), sdp.bilinear_bases[2 * j + 1], n);
          json_vector(data.at("c"), sdp.primal_objective_c[j]);
          json_matrix(data.at("B"), sdp.free_var_matrix[j], sdp.N());
        }
      else
        throw std::runtime_error("Missing block data: neither " + bin + " nor " + sdp_dir + "/block_data" + suffix);
      // SDP::validate (SDP.hxx / SDP/SDP.cxx)
      for(int p = 0; p < 2; ++p)
        if(sdp.bilinear_bases[2 * j + p].h != block_info.bilinear_bases_height(j, p)
           || (sdp.bilinear_bases[2 * j + p].h && sdp.bilinear_bases[2 * j + p].w != n))
          throw std::runtime_error("block " + std::to_string(j) + ": bilinear basis has the wrong size");
      if(sdp.primal_objective_c[j].h != block_info.schur_block_size(j)
         || sdp.free_var_matrix[j].h != block_info.schur_block_size(j)
         || (sdp.free_var_matrix[j].h && sdp.free_var_matrix[j].w != sdp.N()))
        throw std::runtime_error("block " + std::to_string(j) + ": c / B have the wrong size");
    }
}
// The same SDP with every block_data_<j>.json rewritten as block_data_<j>.bin (pmp2sdp
// --outputFormat=bin, src/pmp2sdp/write_block_data.cxx); control / objectives / block_info /
// normalization files are copied.
inline void convert_sdp_to_binary(const std::string &in_dir, const std::string &out_dir)
{
  Block_Info block_info;
  SDP sdp;
  read_sdp(in_dir, block_info, sdp);
  create_directories(out_dir);
  auto copy = [&](const std::string &name) {
    if(!file_exists(in_dir + "/" + name))
      return;
    std::ifstream src(in_dir + "/" + name, std::ios::binary);
    std::ofstream dst(out_dir + "/" + name, std::ios::binary);
    dst << src.rdbuf();
  };
  for(const char *name : {"control.json", "objectives.json", "normalization.json", "pmp_info.json"})
    copy(name);
  for(int j = 0; j < block_info.num_blocks(); ++j)
    {
      copy("block_info_" + std::to_string(j) + ".json");
      write_block_data_bin(out_dir + "/block_data_" + std::to_string(j) + ".bin", sdp.free_var_matrix[j],
                           sdp.primal_objective_c[j], sdp.bilinear_bases[2 * j], sdp.bilinear_bases[2 * j + 1]);
    }
}
} // namespace sdpb_host

// SDPs packed by `pmp2sdp --zip`: a zip archive whose entries are STORED, never deflated
// (reference src/pmp2sdp/Archive_Writer.cxx:10-14, "Hard code zip with no compression"; read back
// through libarchive at src/sdpb_util/Archive_Reader.cxx).  libarchive is not in this image and a
// stored zip needs no decompressor: this reads the central directory (with its zip64 extensions --
// libarchive streams, so local headers carry data descriptors and the sizes live in the directory)
// and copies every entry out.  The reference's own test/data/sdp.zip is the fixture
// (tests/golden/sdp.zip).
#pragma once
#include "serialize.hpp"

#include <cstdlib>
#include <dirent.h>
#include <unistd.h>

namespace sdpb_host
{
inline bool is_regular_file(const std::string &p)
{
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
inline uint64_t le(const std::vector<uint8_t> &b, size_t at, int bytes)
{
  if(at + bytes > b.size())
    throw std::runtime_error("zip: truncated archive");
  uint64_t v = 0;
  for(int i = bytes - 1; i >= 0; --i)
    v = (v << 8) | b[at + i];
  return v;
}
// Extracts every (stored) entry of `zip_path` into `out_dir`; entries in sub-directories keep their path.
inline void extract_stored_zip(const std::string &zip_path, const std::string &out_dir)
{
  std::ifstream in(zip_path, std::ios::binary);
  if(!in)
    throw std::runtime_error("Unable to open: " + zip_path);
  std::vector<uint8_t> b((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  // end of central directory record: signature 0x06054b50 within the last 64 KB + 22 bytes
  size_t eocd = std::string::npos;
  for(size_t at = b.size() >= 22 ? b.size() - 22 : 0;; --at)
    {
      if(b.size() >= 22 && le(b, at, 4) == 0x06054b50u)
        {
          eocd = at;
          break;
        }
      if(at == 0 || b.size() - at > 65557)
        break;
    }
  if(eocd == std::string::npos)
    throw std::runtime_error("Not a zip archive: " + zip_path);
  uint64_t count = le(b, eocd + 10, 2), cd_off = le(b, eocd + 16, 4);
  if(count == 0xFFFF || cd_off == 0xFFFFFFFFu)
    {
      // zip64: locator (0x07064b50) right before the EOCD points at the zip64 EOCD (0x06064b50)
      if(eocd < 20 || le(b, eocd - 20, 4) != 0x07064b50u)
        throw std::runtime_error("zip: missing zip64 locator in " + zip_path);
      const uint64_t e64 = le(b, eocd - 20 + 8, 8);
      if(le(b, e64, 4) != 0x06064b50u)
        throw std::runtime_error("zip: bad zip64 end record in " + zip_path);
      count = le(b, e64 + 32, 8);
      cd_off = le(b, e64 + 48, 8);
    }
  create_directories(out_dir);
  size_t at = cd_off;
  for(uint64_t e = 0; e < count; ++e)
    {
      if(le(b, at, 4) != 0x02014b50u)
        throw std::runtime_error("zip: bad central directory entry in " + zip_path);
      const uint64_t method = le(b, at + 10, 2);
      uint64_t csize = le(b, at + 20, 4), usize = le(b, at + 24, 4), lho = le(b, at + 42, 4);
      const size_t nlen = le(b, at + 28, 2), xlen = le(b, at + 30, 2), clen = le(b, at + 32, 2);
      const std::string name(reinterpret_cast<const char *>(b.data()) + at + 46, nlen);
      // zip64 extended information (header id 1): the fields that read 0xFFFFFFFF, in this order
      for(size_t x = at + 46 + nlen; x + 4 <= at + 46 + nlen + xlen;)
        {
          const uint64_t id = le(b, x, 2), sz = le(b, x + 2, 2);
          if(id == 1)
            {
              size_t f = x + 4;
              if(usize == 0xFFFFFFFFu)
                {
                  usize = le(b, f, 8);
                  f += 8;
                }
              if(csize == 0xFFFFFFFFu)
                {
                  csize = le(b, f, 8);
                  f += 8;
                }
              if(lho == 0xFFFFFFFFu)
                lho = le(b, f, 8);
            }
          x += 4 + sz;
        }
      at += 46 + nlen + xlen + clen;
      if(name.empty() || name.back() == '/')
        {
          create_directories(out_dir + "/" + name);
          continue;
        }
      if(method != 0 || csize != usize)
        throw std::runtime_error("zip: entry '" + name + "' of " + zip_path
                                 + " is compressed; pmp2sdp writes stored entries (Archive_Writer.cxx:10-14)");
      if(name.find("..") != std::string::npos)
        throw std::runtime_error("zip: refusing entry '" + name + "'");
      if(le(b, lho, 4) != 0x04034b50u)
        throw std::runtime_error("zip: bad local header of '" + name + "' in " + zip_path);
      const size_t data = lho + 30 + le(b, lho + 26, 2) + le(b, lho + 28, 2);
      if(data + usize > b.size())
        throw std::runtime_error("zip: truncated entry '" + name + "' in " + zip_path);
      const size_t slash = name.rfind('/');
      if(slash != std::string::npos)
        create_directories(out_dir + "/" + name.substr(0, slash));
      std::ofstream out(out_dir + "/" + name, std::ios::binary);
      out.write(reinterpret_cast<const char *>(b.data()) + data, (std::streamsize)usize);
      if(!out.good())
        throw std::runtime_error("Error when writing to: " + out_dir + "/" + name);
    }
}
// RAII: a zip extracted into a fresh directory under $TMPDIR, removed again on destruction
struct Extracted_Zip
{
  std::string dir;
  explicit Extracted_Zip(const std::string &zip_path)
  {
    const char *tmp = getenv("TMPDIR");
    std::string templ = std::string(tmp && *tmp ? tmp : "/tmp") + "/sdpb_b200_sdp_XXXXXX";
    std::vector<char> buf(templ.begin(), templ.end());
    buf.push_back(0);
    if(!mkdtemp(buf.data()))
      throw std::runtime_error("Unable to create a temporary directory for " + zip_path);
    dir = buf.data();
    try
      {
        extract_stored_zip(zip_path, dir);
      }
    catch(...)
      {
        remove_tree(dir); // a throwing constructor runs no destructor
        throw;
      }
  }
  static void remove_tree(const std::string &p)
  {
    if(DIR *d = opendir(p.c_str()))
      {
        while(struct dirent *e = readdir(d))
          {
            const std::string n = e->d_name;
            if(n == "." || n == "..")
              continue;
            const std::string q = p + "/" + n;
            if(is_directory(q))
              remove_tree(q);
            else
              unlink(q.c_str());
          }
        closedir(d);
      }
    rmdir(p.c_str());
  }
  ~Extracted_Zip()
  {
    if(!dir.empty())
      remove_tree(dir);
  }
};
} // namespace sdpb_host

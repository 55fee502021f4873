// Command-line conventions of the reference drivers, re-implemented for the B200 host programs.
// ref: struct argsparse, cpu/include/argparse.hpp:14-230 (used by gpu/src/LICH-train_rbm.cu:41).  Same observable behaviour:
//   * arguments are "-name=value"; an argument is matched to the FIRST declared option whose name equals the characters
//     following the dash (prefix match in declaration order -- so "-nh=.." is tested against "L", "nh", ... in turn);
//   * "--help" prints the option table (with defaults) and exits 1; a missing option, a doubled option, a missing '=' or an
//     empty value print the reference's "# error(in..)" lines and exit 1;
//   * find<T>() lexical-casts through a stringstream, mfind<T>() splits comma lists; print() echoes the table.
#pragma once
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace nqs_host
{
using pair_t = std::pair<std::string, std::string>;

class argsparse
{
public:
  argsparse(const int argc, const char * const argv[], const std::vector<pair_t> & options,
    const std::vector<pair_t> & defaults = std::vector<pair_t>())
  {
    for (const auto & o : options) { names_.push_back(o.first); values_.emplace_back(); given_.push_back(false); }
    std::vector<std::string> words;
    for (int i = 1; i < argc; ++i) words.emplace_back(argv[i]);
    for (const auto & wd : words)
      if (wd == "--help") help_and_exit(argv[0], options, defaults);
    bool bad = false;
    for (const auto & wd : words)
    {
      for (size_t i = 0; i < names_.size(); ++i)
      {
        const std::string & nm = names_[i];
        if (wd.size() < 1 || wd.compare(1, nm.size(), nm) != 0) continue;
        if (wd.size() <= nm.size()+2)
        { std::cerr << "# error(in-1) ---> Put the option correctly! : " << nm << std::endl; bad = true; break; }
        if (given_[i])
        { std::cerr << "# error(in-2) ---> The doubly occupied option is found! : " << nm << ": " << wd << std::endl; bad = true; break; }
        if (wd[1+nm.size()] != '=')
        { std::cerr << "# error(in-3) ---> The symbol '=' must be in between the option and the argument. : " << wd << std::endl; bad = true; break; }
        values_[i] = wd.substr(nm.size()+2);
        given_[i] = true;
        break;
      }
    }
    for (size_t i = 0; i < names_.size(); ++i)
    {
      if (given_[i]) continue;
      for (const auto & d : defaults)
        if (d.first == names_[i]) { values_[i] = d.second; given_[i] = true; break; }
    }
    for (size_t i = 0; i < names_.size(); ++i)
      if (!given_[i])
      { std::cerr << "# error(in) ---> The following option is missing. : " << names_[i] << std::endl; bad = true; }
    if (bad)
    {
      std::cerr << std::endl << " (hint) Type the command to the cmd line as follows: " << argv[0] << " --help" << std::endl;
      std::exit(1);
    }
  }

  template <typename T = std::string>
  T find(const std::string & name) const { return cast<T>(raw(name), raw(name), ""); }

  template <typename T = std::string>
  std::vector<T> mfind(const std::string & name) const
  {
    const std::string all = raw(name);
    std::vector<T> out;
    size_t b = 0;
    for (;;)
    {
      const size_t e = all.find(',', b);
      const std::string item = all.substr(b, e == std::string::npos ? std::string::npos : e-b);
      if (item.empty()) { std::cerr << "# error has occured: remove ',' at the last part" << std::endl; std::exit(1); }
      out.push_back(cast<T>(item, all, " (option: "+name+")"));
      if (e == std::string::npos) break;
      b = e+1;
    }
    return out;
  }

  template <typename Stream>
  void print(Stream & os) const
  {
    os << "#===== updated arguments =====" << std::endl;
    for (size_t i = 0; i < names_.size(); ++i)
      os << "# " << std::setw(8) << names_[i] << " : " << values_[i] << std::endl;
    os << "#=============================" << std::endl;
  }

private:
  std::string raw(const std::string & name) const
  {
    for (size_t i = 0; i < names_.size(); ++i)
      if (names_[i] == name) return values_[i];
    std::cerr << "# error(out) ---> Threre is no option for your calling. : " << name << std::endl;
    std::exit(1);
  }
  template <typename T>
  static T cast(const std::string & item, const std::string & shown, const std::string & tail)
  {
    std::stringstream ss;
    T v;
    ss << item;
    ss >> v;
    if (ss.fail()) { std::cerr << "# error has occured in the lexical cast: " << shown << tail << std::endl; std::exit(1); }
    return v;
  }
  static void help_and_exit(const char * prog, const std::vector<pair_t> & options, const std::vector<pair_t> & defaults)
  {
    std::cout << " # option list" << std::endl;
    for (const auto & o : options)
    {
      std::cout << std::setw(8) << o.first << " : " << o.second;
      for (const auto & d : defaults)
        if (d.first == o.first) std::cout << " (default : " << d.second << ")";
      std::cout << std::endl;
    }
    std::cout << std::endl << " (hint) " << prog << " -option1=value1 -option2=value2 ..." << std::endl;
    std::exit(1);
  }
  std::vector<std::string> names_, values_;
  std::vector<bool> given_;
};
} // namespace nqs_host

// ref_shim: read_xml for the boost-serialization calibration files (test infrastructure).
#pragma once
#include <cctype>
#include <fstream>
#include <sstream>
#include <string>
#include <boost/property_tree/ptree.hpp>
namespace boost { namespace property_tree {
namespace xml_parser {
static const int trim_whitespace = 4;
struct xml_parser_error : boost::exception {};
namespace detail {
inline std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}
inline void parse_children(const std::string& x, size_t& i, ptree& node, const std::string& closing) {
  std::string text;
  while (i < x.size()) {
    if (x[i] != '<') { text += x[i++]; continue; }
    if (x.compare(i, 4, "<!--") == 0) { i = x.find("-->", i); if (i == std::string::npos) throw xml_parser_error(); i += 3; continue; }
    if (x.compare(i, 2, "<?") == 0) { i = x.find("?>", i); if (i == std::string::npos) throw xml_parser_error(); i += 2; continue; }
    if (x.compare(i, 2, "<!") == 0) { i = x.find('>', i); if (i == std::string::npos) throw xml_parser_error(); i += 1; continue; }
    if (x.compare(i, 2, "</") == 0) {
      size_t e = x.find('>', i);
      if (e == std::string::npos || trim(x.substr(i + 2, e - i - 2)) != closing) throw xml_parser_error();
      i = e + 1;
      node.data() = trim(text);
      return;
    }
    size_t e = x.find('>', i);
    if (e == std::string::npos) throw xml_parser_error();
    std::string tag = x.substr(i + 1, e - i - 1);
    const bool self = !tag.empty() && tag[tag.size() - 1] == '/';
    if (self) tag.resize(tag.size() - 1);
    size_t sp = 0;
    while (sp < tag.size() && !std::isspace((unsigned char)tag[sp])) ++sp;
    const std::string name = tag.substr(0, sp);
    i = e + 1;
    ptree child;
    if (!self) parse_children(x, i, child, name);
    node.add_child(name, child);
  }
  if (!closing.empty()) throw xml_parser_error();
  node.data() = trim(text);
}
}  // namespace detail
}  // namespace xml_parser
inline void read_xml(const std::string& filename, ptree& pt, int = 0) {
  std::ifstream f(filename.c_str());
  if (!f) throw xml_parser::xml_parser_error();
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string x = ss.str();
  size_t i = 0;
  pt = ptree();
  xml_parser::detail::parse_children(x, i, pt, "");
}
} }

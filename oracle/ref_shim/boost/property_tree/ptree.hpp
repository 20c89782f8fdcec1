// ref_shim: a small element tree with the boost::property_tree::ptree calls that
// HDLParser::loadCorrectionsFile makes (test infrastructure).
#pragma once
#include <string>
#include <utility>
#include <vector>
#include <boost/shared_ptr.hpp>
namespace boost { namespace property_tree {
struct ptree_error : boost::exception {};
struct ptree_bad_path : ptree_error {};
class ptree {
 public:
  typedef std::pair<const std::string, ptree> value_type;
  typedef std::vector<value_type>::iterator iterator;
  typedef std::vector<value_type>::const_iterator const_iterator;
  ptree() {}
  ptree(const ptree& o) : data_(o.data_), kids_(o.kids_) {}
  ptree& operator=(const ptree& o) {
    if (this != &o) {
      data_ = o.data_;
      kids_.clear();
      for (const auto& k : o.kids_) kids_.push_back(k);
    }
    return *this;
  }
  std::string& data() { return data_; }
  const std::string& data() const { return data_; }
  iterator begin() { return kids_.begin(); }
  iterator end() { return kids_.end(); }
  const_iterator begin() const { return kids_.begin(); }
  const_iterator end() const { return kids_.end(); }
  ptree& add_child(const std::string& name, const ptree& c) {
    kids_.push_back(value_type(name, c));
    return kids_.back().second;
  }
  ptree& get_child(const std::string& path) {
    ptree* cur = this;
    size_t pos = 0;
    while (pos <= path.size()) {
      size_t dot = path.find('.', pos);
      if (dot == std::string::npos) dot = path.size();
      const std::string key = path.substr(pos, dot - pos);
      ptree* next = nullptr;
      for (auto& k : cur->kids_)
        if (k.first == key) { next = &k.second; break; }
      if (!next) throw ptree_bad_path();
      cur = next;
      pos = dot + 1;
    }
    return *cur;
  }
 private:
  std::string data_;
  std::vector<value_type> kids_;
};
} }

// TEST INFRASTRUCTURE — CPU oracle. Not part of the shipped product path.
// Minimal S-expression reader for the plan/expression text the tests feed to both
// the oracle and (through its own, separate reader) the product library.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

struct SNode {
  bool is_list = false;
  bool quoted = false;  // atom was a "string literal"
  std::string atom;
  std::vector<SNode> kids;
  const std::string& head() const {
    if (!is_list || kids.empty() || kids[0].is_list) throw std::runtime_error("sexpr: expected (head ...)");
    return kids[0].atom;
  }
  size_t nargs() const { return kids.size() - 1; }
  const SNode& arg(size_t i) const {
    if (i + 1 >= kids.size()) throw std::runtime_error("sexpr: missing argument in (" + head() + ")");
    return kids[i + 1];
  }
};

class SParser {
 public:
  explicit SParser(const std::string& s) : s_(s) {}
  SNode parse() {
    SNode n = node();
    ws();
    if (p_ != s_.size()) throw std::runtime_error("sexpr: trailing input");
    return n;
  }

 private:
  void ws() { while (p_ < s_.size() && isspace(static_cast<unsigned char>(s_[p_]))) ++p_; }
  SNode node() {
    ws();
    if (p_ >= s_.size()) throw std::runtime_error("sexpr: unexpected end");
    SNode n;
    if (s_[p_] == '(') {
      ++p_;
      n.is_list = true;
      for (;;) {
        ws();
        if (p_ >= s_.size()) throw std::runtime_error("sexpr: unbalanced (");
        if (s_[p_] == ')') { ++p_; break; }
        n.kids.push_back(node());
      }
    } else if (s_[p_] == '"') {
      ++p_;
      n.quoted = true;
      while (p_ < s_.size() && s_[p_] != '"') {
        if (s_[p_] == '\\' && p_ + 1 < s_.size()) ++p_;
        n.atom.push_back(s_[p_++]);
      }
      if (p_ >= s_.size()) throw std::runtime_error("sexpr: unterminated string");
      ++p_;
    } else {
      while (p_ < s_.size() && !isspace(static_cast<unsigned char>(s_[p_])) && s_[p_] != '(' && s_[p_] != ')')
        n.atom.push_back(s_[p_++]);
    }
    return n;
  }
  const std::string& s_;
  size_t p_ = 0;
};

}  // namespace orc

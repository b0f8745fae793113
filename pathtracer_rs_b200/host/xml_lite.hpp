// Minimal XML reader for Mitsuba 0.5 scene files (the subset quick_xml + serde see in
// src/common/importer/mitsuba.rs): elements, attributes, nesting, comments, the <?xml ?> prolog and the
// five predefined entities.  No DTDs, namespaces or CDATA.  Throws std::runtime_error with a byte offset.
#pragma once
#include <cctype>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ptrs_host {

struct XmlNode {
  std::string name;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<std::unique_ptr<XmlNode>> children;
  std::string text;

  const std::string* attr(const std::string& key) const {
    for (const auto& a : attrs)
      if (a.first == key) return &a.second;
    return nullptr;
  }
  std::string attr_or(const std::string& key, const std::string& dflt) const {
    const std::string* v = attr(key);
    return v ? *v : dflt;
  }
  const XmlNode* child(const std::string& tag) const {
    for (const auto& c : children)
      if (c->name == tag) return c.get();
    return nullptr;
  }
  std::vector<const XmlNode*> all(const std::string& tag) const {
    std::vector<const XmlNode*> out;
    for (const auto& c : children)
      if (c->name == tag) out.push_back(c.get());
    return out;
  }
};

class XmlParser {
 public:
  explicit XmlParser(const std::string& src) : s_(src) {}
  std::unique_ptr<XmlNode> parse_document() {
    skip_misc();
    std::unique_ptr<XmlNode> root = parse_element();
    skip_misc();
    if (i_ != s_.size()) fail("content after the root element");
    return root;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;

  [[noreturn]] void fail(const std::string& what) const { throw std::runtime_error("XML: " + what + " at byte " + std::to_string(i_)); }
  bool starts(const char* lit) const { return s_.compare(i_, std::char_traits<char>::length(lit), lit) == 0; }
  void skip_ws() {
    while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\t' || s_[i_] == '\n' || s_[i_] == '\r')) ++i_;
  }
  void skip_until(const char* lit) {
    const size_t p = s_.find(lit, i_);
    if (p == std::string::npos) fail(std::string("unterminated construct, expected ") + lit);
    i_ = p + std::char_traits<char>::length(lit);
  }
  void skip_misc() {  // whitespace, comments, processing instructions, doctype
    for (;;) {
      skip_ws();
      if (starts("<!--")) skip_until("-->");
      else if (starts("<?")) skip_until("?>");
      else if (starts("<!DOCTYPE")) skip_until(">");
      else return;
    }
  }
  static bool name_char(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }
  std::string parse_name() {
    const size_t b = i_;
    while (i_ < s_.size() && name_char(s_[i_])) ++i_;
    if (i_ == b) fail("expected a name");
    return s_.substr(b, i_ - b);
  }
  static std::string unescape(const std::string& v) {
    if (v.find('&') == std::string::npos) return v;
    static const std::pair<const char*, char> ents[] = {{"&lt;", '<'}, {"&gt;", '>'}, {"&amp;", '&'}, {"&quot;", '"'}, {"&apos;", '\''}};
    std::string out;
    for (size_t k = 0; k < v.size();) {
      bool hit = false;
      if (v[k] == '&')
        for (const auto& e : ents) {
          const size_t n = std::char_traits<char>::length(e.first);
          if (v.compare(k, n, e.first) == 0) {
            out.push_back(e.second);
            k += n;
            hit = true;
            break;
          }
        }
      if (!hit) out.push_back(v[k++]);
    }
    return out;
  }
  std::unique_ptr<XmlNode> parse_element() {
    if (i_ >= s_.size() || s_[i_] != '<') fail("expected '<'");
    ++i_;
    auto node = std::make_unique<XmlNode>();
    node->name = parse_name();
    for (;;) {
      skip_ws();
      if (i_ >= s_.size()) fail("unterminated start tag");
      if (s_[i_] == '/') {
        if (!starts("/>")) fail("expected '/>'");
        i_ += 2;
        return node;
      }
      if (s_[i_] == '>') {
        ++i_;
        break;
      }
      std::string key = parse_name();
      skip_ws();
      if (i_ >= s_.size() || s_[i_] != '=') fail("expected '=' after attribute name");
      ++i_;
      skip_ws();
      if (i_ >= s_.size() || (s_[i_] != '"' && s_[i_] != '\'')) fail("expected a quoted attribute value");
      const char q = s_[i_++];
      const size_t e = s_.find(q, i_);
      if (e == std::string::npos) fail("unterminated attribute value");
      node->attrs.emplace_back(std::move(key), unescape(s_.substr(i_, e - i_)));
      i_ = e + 1;
    }
    for (;;) {  // content
      const size_t lt = s_.find('<', i_);
      if (lt == std::string::npos) fail("unterminated element <" + node->name + ">");
      node->text += unescape(s_.substr(i_, lt - i_));
      i_ = lt;
      if (starts("<!--")) {
        skip_until("-->");
      } else if (starts("<?")) {
        skip_until("?>");
      } else if (starts("</")) {
        i_ += 2;
        const std::string close = parse_name();
        if (close != node->name) fail("mismatched </" + close + "> for <" + node->name + ">");
        skip_ws();
        if (i_ >= s_.size() || s_[i_] != '>') fail("expected '>'");
        ++i_;
        return node;
      } else {
        node->children.push_back(parse_element());
      }
    }
  }
};

}  // namespace ptrs_host

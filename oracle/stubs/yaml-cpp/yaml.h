// oracle/stubs/yaml-cpp/yaml.h -- TEST INFRASTRUCTURE ONLY: a reader for the YAML subset of aLENS' RunConfig.yaml files
// (block mappings, flow sequences [a, b, c], block sequences of mappings for `boundaries:`, comments) behind the
// yaml-cpp names SimToolbox uses: YAML::LoadFile, Node::operator[], as<T>(), size(), iteration, operator bool.
#pragma once
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace YAML {
class Node {
  public:
    enum Kind { Undefined, Scalar, Sequence, Map };
    Node() : d_(std::make_shared<Data>()) {}
    explicit operator bool() const { return d_->kind != Undefined; }
    bool operator!() const { return d_->kind == Undefined; }
    bool IsDefined() const { return d_->kind != Undefined; }
    bool IsSequence() const { return d_->kind == Sequence; }
    bool IsMap() const { return d_->kind == Map; }
    bool IsScalar() const { return d_->kind == Scalar; }
    size_t size() const { return d_->kind == Sequence ? d_->seq.size() : d_->kind == Map ? d_->map.size() : 0; }
    Node operator[](const std::string &k) const {
        auto it = d_->map.find(k);
        return it == d_->map.end() ? Node() : it->second;
    }
    Node operator[](const char *k) const { return (*this)[std::string(k)]; }
    Node operator[](int i) const { return (d_->kind == Sequence && (size_t)i < d_->seq.size()) ? d_->seq[i] : Node(); }
    Node operator[](size_t i) const { return (*this)[(int)i]; }
    std::vector<Node>::const_iterator begin() const { return d_->seq.begin(); }
    std::vector<Node>::const_iterator end() const { return d_->seq.end(); }
    const std::string &Scalar_() const { return d_->scalar; }
    template <class T>
    T as() const {
        if (d_->kind != Scalar) throw std::runtime_error("yaml stub: not a scalar");
        return convert<T>(d_->scalar);
    }
    // construction (parser)
    void setScalar(const std::string &s) { d_->kind = Scalar; d_->scalar = s; }
    void push(const Node &n) { d_->kind = Sequence; d_->seq.push_back(n); }
    void set(const std::string &k, const Node &n) { d_->kind = Map; d_->map[k] = n; }

  private:
    struct Data {
        Kind kind = Undefined;
        std::string scalar;
        std::vector<Node> seq;
        std::map<std::string, Node> map;
    };
    std::shared_ptr<Data> d_;
    template <class T>
    static T convert(const std::string &s) {
        std::istringstream is(s);
        T v;
        is >> v;
        if (is.fail()) throw std::runtime_error("yaml stub: bad conversion of '" + s + "'");
        return v;
    }
};
template <>
inline std::string Node::convert<std::string>(const std::string &s) { return s; }
template <>
inline bool Node::convert<bool>(const std::string &s) {
    if (s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "on" || s == "1") return true;
    if (s == "false" || s == "False" || s == "FALSE" || s == "no" || s == "off" || s == "0") return false;
    throw std::runtime_error("yaml stub: bad bool '" + s + "'");
}

namespace detail {
inline std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
inline std::string unquote(std::string s) {
    s = trim(s);
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) s = s.substr(1, s.size() - 2);
    return s;
}
inline Node value(const std::string &raw) {
    const std::string v = trim(raw);
    Node n;
    if (!v.empty() && v.front() == '[') {
        const size_t e = v.rfind(']');
        std::string body = v.substr(1, (e == std::string::npos ? v.size() : e) - 1);
        std::stringstream ss(body);
        std::string item;
        Node seq;
        while (std::getline(ss, item, ',')) {
            Node s;
            s.setScalar(unquote(item));
            seq.push(s);
        }
        if (!seq.IsSequence()) { // "[]"
            Node dummy; seq.push(dummy); seq = Node(); seq.push(dummy);
        }
        return seq;
    }
    n.setScalar(unquote(v));
    return n;
}
struct Line {
    int indent;
    std::string text;
};
// parse lines[i...) that are indented deeper than `parent` into a mapping or a sequence
inline Node block(const std::vector<Line> &L, size_t &i, int parent) {
    Node out;
    if (i >= L.size() || L[i].indent <= parent) return out;
    const int ind = L[i].indent;
    while (i < L.size() && L[i].indent == ind) {
        std::string t = L[i].text;
        if (t.rfind("- ", 0) == 0 || t == "-") { // sequence item; its mapping continues on deeper lines
            std::vector<Line> sub;
            std::string first = t.size() > 2 ? t.substr(2) : "";
            const int subInd = ind + 2;
            if (!trim(first).empty()) sub.push_back(Line{subInd, trim(first)});
            i++;
            while (i < L.size() && L[i].indent > ind) sub.push_back(L[i++]);
            for (auto &s : sub) if (s.indent < subInd) s.indent = subInd;
            size_t j = 0;
            Node item = (sub.size() == 1 && sub[0].text.find(':') == std::string::npos) ? value(sub[0].text) : block(sub, j, subInd - 1);
            out.push(item);
            continue;
        }
        const size_t c = t.find(':');
        if (c == std::string::npos) throw std::runtime_error("yaml stub: cannot parse line '" + t + "'");
        const std::string key = unquote(t.substr(0, c)), rest = trim(t.substr(c + 1));
        i++;
        if (rest.empty()) out.set(key, block(L, i, ind));
        else out.set(key, value(rest));
    }
    return out;
}
} // namespace detail

inline Node Load(std::istream &in) {
    std::vector<detail::Line> L;
    std::string line;
    while (std::getline(in, line)) {
        bool q = false;
        for (size_t k = 0; k < line.size(); k++) { // strip comments outside quotes
            if (line[k] == '"' || line[k] == '\'') q = !q;
            if (line[k] == '#' && !q && (k == 0 || line[k - 1] == ' ' || line[k - 1] == '\t')) { line = line.substr(0, k); break; }
        }
        const size_t a = line.find_first_not_of(" \t");
        if (a == std::string::npos) continue;
        const std::string t = detail::trim(line);
        if (t == "---" || t == "...") continue;
        L.push_back(detail::Line{(int)a, t});
    }
    size_t i = 0;
    return detail::block(L, i, -1);
}
inline Node LoadFile(const std::string &fn) {
    std::ifstream f(fn);
    if (!f) throw std::runtime_error("yaml stub: cannot open " + fn);
    return Load(f);
}
} // namespace YAML

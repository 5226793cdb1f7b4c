// YAML-subset reader, see rgh_yaml.h.
#include "rgh_yaml.h"

#include <cctype>
#include <cerrno>
#include <cmath>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "rgh_internal.h"

namespace rgh {
namespace {

struct Line {
    int indent;
    std::string text;  // after the indentation, comment stripped, right-trimmed; never empty
    int no;            // 1-based source line
};

struct ParseError {
    int code;
    int line;
    std::string msg;
};

inline bool is_space(char c) { return c == ' ' || c == '\t'; }

void rtrim(std::string &s) {
    while (!s.empty() && is_space(s.back())) s.pop_back();
}

// Cuts a trailing comment (a '#' at the start or after white space, outside quotes).
std::string strip_comment(const std::string &s) {
    char q = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (q == '"') {
            if (c == '\\') ++i;
            else if (c == '"') q = 0;
        } else if (q == '\'') {
            if (c == '\'') {
                if (i + 1 < s.size() && s[i + 1] == '\'') ++i;
                else q = 0;
            }
        } else if ((c == '"' || c == '\'') && (i == 0 || is_space(s[i - 1]) || std::strchr("[{,:-", s[i - 1]))) {
            q = c;
        } else if (c == '#' && (i == 0 || is_space(s[i - 1]))) {
            return s.substr(0, i);
        }
    }
    return s;
}

void append_utf8(std::string &out, unsigned cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) {
        out += (char)(0xC0 | (cp >> 6));
        out += (char)(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
        out += (char)(0xE0 | (cp >> 12));
        out += (char)(0x80 | ((cp >> 6) & 0x3F));
        out += (char)(0x80 | (cp & 0x3F));
    } else {
        out += (char)(0xF0 | (cp >> 18));
        out += (char)(0x80 | ((cp >> 12) & 0x3F));
        out += (char)(0x80 | ((cp >> 6) & 0x3F));
        out += (char)(0x80 | (cp & 0x3F));
    }
}

struct Parser {
    std::vector<Line> lines;
    size_t cur = 0;
    std::vector<std::pair<std::string, YamlNode>> anchors;   // &name -> node (a later definition shadows an earlier one)
    int depth = 0;   // nesting of collections being parsed (bounded: the parser recurses)
    struct Nest {
        Parser &p;
        Nest(Parser &q, int line) : p(q) {
            if (++p.depth > 200) p.fail(line, "collections nested deeper than 200 levels", RGH_E_UNSUPPORTED);
        }
        ~Nest() { --p.depth; }
    };

    // "&name rest" -> name, text := rest (left-trimmed).  Anchors name a node for later aliases.
    bool take_anchor(std::string &text, std::string &name, int line) {
        if (text.empty() || text[0] != '&') return false;
        size_t e = 1;
        while (e < text.size() && !is_space(text[e]) && !std::strchr(",[]{}", text[e])) ++e;
        name = text.substr(1, e - 1);
        if (name.empty()) fail(line, "empty anchor name");
        while (e < text.size() && is_space(text[e])) ++e;
        text = text.substr(e);
        return true;
    }
    void define_anchor(const std::string &name, const YamlNode &n) { anchors.emplace_back(name, n); }
    YamlNode alias(const std::string &name, int line) {
        for (size_t i = anchors.size(); i-- > 0;)
            if (anchors[i].first == name) {
                YamlNode n = anchors[i].second;
                return n;
            }
        fail(line, "alias *" + name + " refers to an unknown anchor");
    }

    [[noreturn]] void fail(int line, const std::string &msg, int code = RGH_E_FORMAT) { throw ParseError{code, line, msg}; }

    // ---- scalars -------------------------------------------------------------------------
    // Parses a quoted scalar starting at s[pos] (a quote); leaves pos after the closing quote.
    std::string quoted(const std::string &s, size_t &pos, int line) {
        const char q = s[pos++];
        std::string out;
        for (; pos < s.size(); ++pos) {
            const char c = s[pos];
            if (q == '\'') {
                if (c == '\'') {
                    if (pos + 1 < s.size() && s[pos + 1] == '\'') {
                        out += '\'';
                        ++pos;
                    } else {
                        ++pos;
                        return out;
                    }
                } else {
                    out += c;
                }
            } else {
                if (c == '"') {
                    ++pos;
                    return out;
                }
                if (c != '\\') {
                    out += c;
                    continue;
                }
                if (++pos >= s.size()) break;
                const char e = s[pos];
                auto hex = [&](int n) {
                    unsigned v = 0;
                    for (int i = 0; i < n; ++i) {
                        if (++pos >= s.size() || !std::isxdigit((unsigned char)s[pos])) fail(line, "bad escape in double-quoted scalar");
                        const char h = s[pos];
                        v = v * 16 + (unsigned)(h <= '9' ? h - '0' : (h | 32) - 'a' + 10);
                    }
                    append_utf8(out, v);
                };
                switch (e) {
                    case '0': out += '\0'; break;
                    case 'a': out += '\a'; break;
                    case 'b': out += '\b'; break;
                    case 't': case '\t': out += '\t'; break;
                    case 'n': out += '\n'; break;
                    case 'v': out += '\v'; break;
                    case 'f': out += '\f'; break;
                    case 'r': out += '\r'; break;
                    case 'e': out += '\x1b'; break;
                    case ' ': out += ' '; break;
                    case '"': out += '"'; break;
                    case '/': out += '/'; break;
                    case '\\': out += '\\'; break;
                    case 'N': append_utf8(out, 0x85); break;
                    case '_': append_utf8(out, 0xA0); break;
                    case 'L': append_utf8(out, 0x2028); break;
                    case 'P': append_utf8(out, 0x2029); break;
                    case 'x': hex(2); break;
                    case 'u': hex(4); break;
                    case 'U': hex(8); break;
                    default: fail(line, "unknown escape in double-quoted scalar");
                }
            }
        }
        fail(line, "unterminated quoted scalar (multi-line quoted scalars are not supported)");
    }

    static void reject_unsupported_start(Parser *p, char c, int line) {
        if (c == '!') p->fail(line, "tags are not supported", RGH_E_UNSUPPORTED);
        if (c == '|' || c == '>') p->fail(line, "block scalars are not supported", RGH_E_UNSUPPORTED);
        if (c == '%' || c == '@' || c == '`') p->fail(line, "a plain scalar cannot start with this character");
    }

    YamlNode plain(std::string text, int line) {
        rtrim(text);
        YamlNode n;
        n.line = line;
        if (text.empty()) return n;  // Null
        n.kind = YamlNode::Scalar;
        n.text = text;
        return n;
    }

    // ---- flow collections ------------------------------------------------------------------
    static void skip_ws(const std::string &s, size_t &pos) {
        while (pos < s.size() && (is_space(s[pos]) || s[pos] == '\n')) ++pos;
    }

    YamlNode flow_node(const std::string &s, size_t &pos, int line, bool in_map_value) {
        Nest nest(*this, line);
        skip_ws(s, pos);
        YamlNode n;
        n.line = line;
        if (pos >= s.size()) return n;
        const char c = s[pos];
        if (c == '[') {
            ++pos;
            n.kind = YamlNode::Seq;
            for (;;) {
                skip_ws(s, pos);
                if (pos >= s.size()) fail(line, "unterminated flow sequence");
                if (s[pos] == ']') {
                    ++pos;
                    return n;
                }
                n.items.push_back(flow_node(s, pos, line, false));
                skip_ws(s, pos);
                if (pos < s.size() && s[pos] == ',') {
                    ++pos;
                    continue;
                }
                if (pos < s.size() && s[pos] == ']') continue;
                fail(line, "expected ',' or ']' in flow sequence");
            }
        }
        if (c == '{') {
            ++pos;
            n.kind = YamlNode::Map;
            for (;;) {
                skip_ws(s, pos);
                if (pos >= s.size()) fail(line, "unterminated flow mapping");
                if (s[pos] == '}') {
                    ++pos;
                    return n;
                }
                std::string key;
                if (s[pos] == '"' || s[pos] == '\'') {
                    key = quoted(s, pos, line);
                } else {
                    const size_t start = pos;
                    while (pos < s.size() && s[pos] != ',' && s[pos] != '}' &&
                           !(s[pos] == ':' && (pos + 1 >= s.size() || is_space(s[pos + 1]) || s[pos + 1] == ',' || s[pos + 1] == '}')))
                        ++pos;
                    key = s.substr(start, pos - start);
                    rtrim(key);
                }
                skip_ws(s, pos);
                YamlNode value;
                value.line = line;
                if (pos < s.size() && s[pos] == ':') {
                    ++pos;
                    value = flow_node(s, pos, line, true);
                }
                for (const auto &e : n.entries)
                    if (e.first == key) fail(line, "duplicate key `" + key + "`");
                n.entries.emplace_back(key, std::move(value));
                skip_ws(s, pos);
                if (pos < s.size() && s[pos] == ',') {
                    ++pos;
                    continue;
                }
                if (pos < s.size() && s[pos] == '}') continue;
                fail(line, "expected ',' or '}' in flow mapping");
            }
        }
        if (c == '"' || c == '\'') {
            n.kind = YamlNode::Scalar;
            n.quoted = true;
            n.text = quoted(s, pos, line);
            return n;
        }
        if (c == ',' || c == ']' || c == '}') return n;  // empty -> Null
        if (c == '&') {   // anchored flow node
            size_t e = pos + 1;
            while (e < s.size() && !is_space(s[e]) && s[e] != '\n' && !std::strchr(",[]{}", s[e])) ++e;
            const std::string name = s.substr(pos + 1, e - pos - 1);
            if (name.empty()) fail(line, "empty anchor name");
            pos = e;
            YamlNode v = flow_node(s, pos, line, in_map_value);
            define_anchor(name, v);
            return v;
        }
        if (c == '*') {
            size_t e = pos + 1;
            while (e < s.size() && !is_space(s[e]) && s[e] != '\n' && !std::strchr(",[]{}", s[e])) ++e;
            const std::string name = s.substr(pos + 1, e - pos - 1);
            pos = e;
            return alias(name, line);
        }
        reject_unsupported_start(this, c, line);
        const size_t start = pos;
        while (pos < s.size() && s[pos] != ',' && s[pos] != ']' && s[pos] != '}' && s[pos] != '\n') ++pos;
        (void)in_map_value;
        return plain(s.substr(start, pos - start), line);
    }

    // An inline value: the rest of a line after "key:" or "- " (or a whole line).  Flow collections
    // may continue on the following lines until their brackets balance.
    YamlNode inline_value(std::string text, int line) {
        std::string anchor;
        if (take_anchor(text, anchor, line)) {
            if (text.empty()) fail(line, "an anchor must be followed by a node on this line or by an indented block");
            YamlNode v = inline_value(text, line);
            define_anchor(anchor, v);
            return v;
        }
        if (text[0] == '*') {
            std::string name = text.substr(1);
            rtrim(name);
            return alias(name, line);
        }
        const char c = text[0];
        if (c == '[' || c == '{') {
            auto balanced = [](const std::string &s) {
                int depth = 0;
                char q = 0;
                for (size_t i = 0; i < s.size(); ++i) {
                    const char ch = s[i];
                    if (q == '"') {
                        if (ch == '\\') ++i;
                        else if (ch == '"') q = 0;
                    } else if (q == '\'') {
                        if (ch == '\'') q = 0;
                    } else if (ch == '"' || ch == '\'') q = ch;
                    else if (ch == '[' || ch == '{') ++depth;
                    else if (ch == ']' || ch == '}') --depth;
                }
                return depth <= 0;
            };
            while (!balanced(text)) {
                if (cur + 1 >= lines.size()) fail(line, "unterminated flow collection");
                ++cur;
                text += '\n';
                text += lines[cur].text;
            }
            size_t pos = 0;
            YamlNode n = flow_node(text, pos, line, false);
            skip_ws(text, pos);
            if (pos != text.size()) fail(line, "unexpected characters after flow collection");
            return n;
        }
        if (c == '"' || c == '\'') {
            size_t pos = 0;
            YamlNode n;
            n.kind = YamlNode::Scalar;
            n.quoted = true;
            n.line = line;
            n.text = quoted(text, pos, line);
            while (pos < text.size() && is_space(text[pos])) ++pos;
            if (pos != text.size()) fail(line, "unexpected characters after quoted scalar");
            return n;
        }
        reject_unsupported_start(this, c, line);
        return plain(text, line);
    }

    // ---- block structure -------------------------------------------------------------------
    static bool is_dash(const std::string &t) { return t[0] == '-' && (t.size() == 1 || is_space(t[1])); }

    // If `t` is "key: rest" / "key:", returns true and splits it.
    bool split_key(const std::string &t, std::string &key, std::string &rest, int line) {
        size_t pos = 0;
        if (t[0] == '"' || t[0] == '\'') {
            key = quoted(t, pos, line);
            while (pos < t.size() && is_space(t[pos])) ++pos;
            if (pos >= t.size() || t[pos] != ':' || (pos + 1 < t.size() && !is_space(t[pos + 1]))) return false;
        } else {
            if (t[0] == '[' || t[0] == '{') return false;
            for (;; ++pos) {
                if (pos >= t.size()) return false;
                if (t[pos] == ':' && (pos + 1 == t.size() || is_space(t[pos + 1]))) break;
            }
            key = t.substr(0, pos);
            rtrim(key);
            if (key.empty()) fail(line, "empty mapping key");
            if (key[0] == '?') fail(line, "complex mapping keys are not supported", RGH_E_UNSUPPORTED);
        }
        ++pos;
        while (pos < t.size() && is_space(t[pos])) ++pos;
        rest = t.substr(pos);
        return true;
    }

    YamlNode block(int indent) {
        Nest nest(*this, lines[cur].no);
        const Line &l = lines[cur];
        if (is_dash(l.text)) return seq(indent);
        std::string key, rest;
        if (split_key(l.text, key, rest, l.no)) return map(indent);
        const int no = l.no;
        YamlNode n = inline_value(l.text, no);
        ++cur;
        if (cur < lines.size() && lines[cur].indent > indent)
            fail(lines[cur].no, "multi-line plain scalars are not supported", RGH_E_UNSUPPORTED);
        return n;
    }

    YamlNode nested_or_null(int parent_indent, int line, bool allow_same_indent_seq) {
        YamlNode n;
        n.line = line;
        if (cur < lines.size()) {
            if (lines[cur].indent > parent_indent) return block(lines[cur].indent);
            if (allow_same_indent_seq && lines[cur].indent == parent_indent && is_dash(lines[cur].text)) return seq(parent_indent);
        }
        return n;
    }

    YamlNode map(int indent) {
        YamlNode n;
        n.kind = YamlNode::Map;
        n.line = lines[cur].no;
        while (cur < lines.size() && lines[cur].indent == indent) {
            const int no = lines[cur].no;
            if (is_dash(lines[cur].text)) fail(no, "sequence entry where a mapping key was expected");
            std::string key, rest;
            if (!split_key(lines[cur].text, key, rest, no)) fail(no, "expected `key: value`");
            for (const auto &e : n.entries)
                if (e.first == key) fail(no, "duplicate key `" + key + "`");
            YamlNode value;
            std::string anchor;
            if (!rest.empty() && rest[0] == '&') {   // "key: &a" with the node on the following lines
                std::string probe = rest, name;
                take_anchor(probe, name, no);
                if (probe.empty()) {
                    anchor = name;
                    rest.clear();
                }
            }
            if (rest.empty()) {
                ++cur;
                value = nested_or_null(indent, no, true);
                if (!anchor.empty()) define_anchor(anchor, value);
            } else {
                value = inline_value(rest, no);
                ++cur;
                if (cur < lines.size() && lines[cur].indent > indent)
                    fail(lines[cur].no, "unexpected indentation after an inline value");
            }
            n.entries.emplace_back(key, std::move(value));
        }
        if (cur < lines.size() && lines[cur].indent > indent) fail(lines[cur].no, "bad indentation of a mapping entry");
        return n;
    }

    YamlNode seq(int indent) {
        YamlNode n;
        n.kind = YamlNode::Seq;
        n.line = lines[cur].no;
        while (cur < lines.size() && lines[cur].indent == indent && is_dash(lines[cur].text)) {
            const int no = lines[cur].no;
            const std::string &t = lines[cur].text;
            size_t pos = 1;
            while (pos < t.size() && is_space(t[pos])) ++pos;
            std::string anchor;
            if (pos < t.size() && t[pos] == '&') {   // "- &a" / "- &a content": the anchor names the whole item
                std::string rest = t.substr(pos), name;
                take_anchor(rest, name, no);
                anchor = name;
                pos = t.size() - rest.size();
            }
            if (pos >= t.size()) {
                ++cur;
                n.items.push_back(nested_or_null(indent, no, false));
                if (!anchor.empty()) define_anchor(anchor, n.items.back());
            } else if (!anchor.empty()) {
                const int inner = indent + (int)pos;
                std::string content = t.substr(pos);
                lines[cur].indent = inner;
                lines[cur].text = content;
                std::string key, rest;
                if (split_key(content, key, rest, no)) {
                    // "- &a key: value": an anchor belongs to the NEXT node, which is the key scalar
                    YamlNode k;
                    k.kind = YamlNode::Scalar;
                    k.text = key;
                    k.line = no;
                    define_anchor(anchor, k);
                    n.items.push_back(block(inner));
                } else {
                    n.items.push_back(block(inner));
                    define_anchor(anchor, n.items.back());
                }
            } else {
                // "- content": re-read `content` as a block starting at its own column
                const int inner = indent + (int)pos;
                std::string content = t.substr(pos);
                lines[cur].indent = inner;
                lines[cur].text = content;
                n.items.push_back(block(inner));
            }
        }
        if (cur < lines.size() && lines[cur].indent > indent) fail(lines[cur].no, "bad indentation of a sequence entry");
        return n;
    }

    void load(const char *text, size_t len) {
        size_t i = 0;
        if (len >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF) i = 3;
        int no = 0;
        bool doc_started = false, doc_ended = false;
        while (i <= len) {
            size_t e = i;
            while (e < len && text[e] != '\n') ++e;
            std::string raw(text + i, e - i);
            i = e + 1;
            ++no;
            if (!raw.empty() && raw.back() == '\r') raw.pop_back();
            if (e >= len && raw.empty()) break;
            if (raw.compare(0, 3, "---") == 0 && (raw.size() == 3 || is_space(raw[3]))) {
                if (doc_started || !lines.empty()) fail(no, "multi-document streams are not supported", RGH_E_UNSUPPORTED);
                doc_started = true;
                raw = raw.size() > 3 ? std::string(4, ' ') + raw.substr(4) : std::string();
            } else if (raw.compare(0, 3, "...") == 0 && (raw.size() == 3 || is_space(raw[3]))) {
                doc_ended = true;
                continue;
            } else if (!raw.empty() && raw[0] == '%' && lines.empty() && !doc_started) {
                continue;  // %YAML / %TAG directive
            }
            std::string s = strip_comment(raw);
            rtrim(s);
            size_t ind = 0;
            while (ind < s.size() && s[ind] == ' ') ++ind;
            if (ind == s.size()) continue;
            if (s[ind] == '\t') fail(no, "tab character used for indentation");
            if (doc_ended) fail(no, "content after the document end marker", RGH_E_UNSUPPORTED);
            lines.push_back(Line{(int)ind, s.substr(ind), no});
        }
    }

    YamlNode run(const char *text, size_t len) {
        load(text, len);
        YamlNode root;
        if (lines.empty()) return root;  // empty document -> Null
        root = block(lines[0].indent);
        if (cur < lines.size()) fail(lines[cur].no, "unexpected content after the document's root node");
        return root;
    }
};

// Rust's `str::parse::<i64>()`: optional sign, decimal digits, no overflow.
bool parse_i64(const std::string &s, long long *out) {
    size_t i = 0;
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) ++i;
    if (i >= s.size()) return false;
    for (size_t k = i; k < s.size(); ++k)
        if (s[k] < '0' || s[k] > '9') return false;
    errno = 0;
    const long long v = std::strtoll(s.c_str(), nullptr, 10);
    if (errno == ERANGE) return false;
    *out = v;
    return true;
}

bool parse_radix(const std::string &digits, int radix, long long *out) {
    size_t i = 0;
    if (i < digits.size() && (digits[i] == '+' || digits[i] == '-')) ++i;
    if (i >= digits.size()) return false;
    for (size_t k = i; k < digits.size(); ++k) {
        const char c = digits[k];
        const int d = c >= '0' && c <= '9' ? c - '0' : ((c | 32) >= 'a' && (c | 32) <= 'f' ? (c | 32) - 'a' + 10 : 99);
        if (d >= radix) return false;
    }
    errno = 0;
    const long long v = std::strtoll(digits.c_str(), nullptr, radix);
    if (errno == ERANGE) return false;
    *out = v;
    return true;
}

// Rust's `str::parse::<f64>()` grammar (2017): [+-] ( "inf" | "NaN" | digits [. digits] [(e|E) [+-] digits] ).
bool parse_f64(const std::string &s, double *out) {
    size_t i = 0;
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) ++i;
    const std::string body = s.substr(i);
    if (body == "inf") {
        *out = s[0] == '-' ? -HUGE_VAL : HUGE_VAL;
        return true;
    }
    if (body == "NaN") {
        *out = std::strtod("nan", nullptr);
        return true;
    }
    size_t k = 0, digits = 0;
    while (k < body.size() && body[k] >= '0' && body[k] <= '9') ++k, ++digits;
    if (k < body.size() && body[k] == '.') {
        ++k;
        while (k < body.size() && body[k] >= '0' && body[k] <= '9') ++k, ++digits;
    }
    if (digits == 0) return false;
    if (k < body.size() && (body[k] == 'e' || body[k] == 'E')) {
        ++k;
        if (k < body.size() && (body[k] == '+' || body[k] == '-')) ++k;
        size_t ed = 0;
        while (k < body.size() && body[k] >= '0' && body[k] <= '9') ++k, ++ed;
        if (ed == 0) return false;
    }
    if (k != body.size()) return false;
    *out = std::strtod(s.c_str(), nullptr);  // correctly rounded (glibc), like Rust's dec2flt
    return true;
}

}  // namespace

int yaml_parse(const char *text, size_t len, YamlNode &root, std::string &error) {
    Parser p;
    try {
        root = p.run(text, len);
    } catch (const ParseError &e) {
        error = "line " + std::to_string(e.line) + ": " + e.msg;
        return e.code;
    }
    return RGH_OK;
}

ScalarType yaml_scalar_type(const YamlNode &n, double *real, long long *integer, bool *boolean) {
    if (n.kind == YamlNode::Null) return ScalarType::Null;
    if (n.kind != YamlNode::Scalar || n.quoted) return ScalarType::String;
    const std::string &v = n.text;
    long long iv = 0;
    double dv = 0;
    if (v.compare(0, 2, "0x") == 0 && parse_radix(v.substr(2), 16, &iv)) {
        *integer = iv;
        return ScalarType::Int;
    }
    if (v.compare(0, 2, "0o") == 0 && parse_radix(v.substr(2), 8, &iv)) {
        *integer = iv;
        return ScalarType::Int;
    }
    if (v[0] == '+' && parse_i64(v.substr(1), &iv)) {
        *integer = iv;
        return ScalarType::Int;
    }
    if (v == "~" || v == "null") return ScalarType::Null;
    if (v == "true" || v == "false") {
        *boolean = v == "true";
        return ScalarType::Bool;
    }
    if (parse_i64(v, &iv)) {
        *integer = iv;
        return ScalarType::Int;
    }
    if (parse_f64(v, &dv)) {
        *real = dv;
        return ScalarType::Real;
    }
    return ScalarType::String;
}

}  // namespace rgh

// A small YAML reader for the scene schema (block + flow collections, plain / quoted scalars,
// comments, one document).  Stands in for yaml-rust 0.3.5 under serde_yaml 0.6.2 (Cargo.lock),
// which the reference uses at src/main.rs:118.  Anchors (&a) and aliases (*a) are resolved by copy;
// tags, block scalars (| >), merge keys and multi-document streams are outside the subset and are
// reported as RGH_E_UNSUPPORTED.
#ifndef RGH_YAML_H
#define RGH_YAML_H

#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace rgh {

struct YamlNode {
    enum Kind { Null, Scalar, Seq, Map } kind = Null;
    std::string text;     // Scalar: the value with quotes / escapes resolved
    bool quoted = false;  // Scalar: written with '...' or "..." (always a string to serde)
    int line = 0;
    std::vector<YamlNode> items;                            // Seq
    std::vector<std::pair<std::string, YamlNode>> entries;  // Map, in document order

    const YamlNode *get(const std::string &key) const {
        for (const auto &e : entries)
            if (e.first == key) return &e.second;
        return nullptr;
    }
};

// Returns 0 or an RGH_E_* code; `error` receives "line N: message".
int yaml_parse(const char *text, size_t len, YamlNode &root, std::string &error);

// Scalar resolution as yaml-rust's Yaml::from_str does it for plain scalars.
enum class ScalarType { Null, Bool, Int, Real, String };
ScalarType yaml_scalar_type(const YamlNode &n, double *real, long long *integer, bool *boolean);

}  // namespace rgh
#endif

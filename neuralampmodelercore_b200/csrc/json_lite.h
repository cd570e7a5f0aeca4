// json_lite.h -- a small recursive-descent JSON reader for .nam files.
//
// The reference reads .nam files with nlohmann::json (third-party, NAM/nam_file.cpp:9-40).
// The product only needs a read-only DOM, so it carries its own ~250-line parser instead of a
// 25 kLoC dependency.  Numbers are kept as double (weights are float in the file anyway,
// NAM/get_dsp.cpp:130-139 converts to std::vector<float>).
#pragma once

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace namb200
{
namespace json
{

/// Shortest round-trip decimal, locale-independent; NaN / Infinity / -Infinity for non-finite values.
std::string number_to_string(double v);


class ParseError : public std::runtime_error
{
public:
  using std::runtime_error::runtime_error;
};

class Value
{
public:
  enum class Type
  {
    Null,
    Bool,
    Number,
    String,
    Array,
    Object
  };

  Value() = default;

  Type type() const { return _type; }
  bool is_null() const { return _type == Type::Null; }
  bool is_bool() const { return _type == Type::Bool; }
  bool is_number() const { return _type == Type::Number; }
  bool is_string() const { return _type == Type::String; }
  bool is_array() const { return _type == Type::Array; }
  bool is_object() const { return _type == Type::Object; }

  // Typed access; throws std::runtime_error naming `what` on a type mismatch.
  bool as_bool(const char* what = "value") const;
  double as_double(const char* what = "value") const;
  int as_int(const char* what = "value") const;
  const std::string& as_string(const char* what = "value") const;
  const std::vector<Value>& items(const char* what = "value") const;

  // Object access
  bool contains(const std::string& key) const;
  // Returns a Null value when the key is absent
  const Value& get(const std::string& key) const;
  // Throws std::runtime_error("missing key ...") when absent
  const Value& at(const std::string& key) const;
  const std::vector<std::pair<std::string, Value>>& members() const { return _members; }

  size_t size() const { return _type == Type::Array ? _items.size() : _members.size(); }
  const Value& operator[](size_t i) const { return _items.at(i); }

  // value-with-default helpers (nlohmann's .value(key, default))
  int value_int(const std::string& key, int dflt) const;
  double value_double(const std::string& key, double dflt) const;
  bool value_bool(const std::string& key, bool dflt) const;

  static Value parse(const std::string& text);
  /// Compact JSON text of this value (numbers printed with enough digits to round-trip a double).
  std::string dump() const;

private:
  friend class Parser;
  Type _type = Type::Null;
  bool _bool = false;
  double _number = 0.0;
  std::string _string;
  std::vector<Value> _items;
  std::vector<std::pair<std::string, Value>> _members; // insertion order kept
};

} // namespace json
} // namespace namb200

#include "json_lite.h"

#include <charconv>
#include <cmath>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace namb200
{
namespace json
{

// Shortest round-trip decimal of a double, locale-independent (std::to_chars); non-finite values as the tokens the parser
// accepts (NaN, Infinity, -Infinity) so that a document survives dump() -> parse().
std::string number_to_string(double v)
{
  if (std::isnan(v))
    return "NaN";
  if (std::isinf(v))
    return v < 0 ? "-Infinity" : "Infinity";
  char buf[64];
  const auto res = std::to_chars(buf, buf + sizeof buf, v);
  return std::string(buf, res.ptr);
}


namespace
{
const Value kNull;
}

class Parser
{
public:
  explicit Parser(const std::string& text)
  : _s(text.data())
  , _n(text.size())
  {
  }

  Value parse_document()
  {
    skip_ws();
    Value v = parse_value(0);
    skip_ws();
    if (_i != _n)
      fail("trailing characters after JSON document");
    return v;
  }

private:
  const char* _s;
  size_t _n;
  size_t _i = 0;

  [[noreturn]] void fail(const std::string& msg) const
  {
    throw ParseError("JSON parse error at byte " + std::to_string(_i) + ": " + msg);
  }

  void skip_ws()
  {
    while (_i < _n && (_s[_i] == ' ' || _s[_i] == '\n' || _s[_i] == '\t' || _s[_i] == '\r'))
      _i++;
  }

  bool consume(const char* lit)
  {
    const size_t len = std::strlen(lit);
    if (_n - _i >= len && std::memcmp(_s + _i, lit, len) == 0)
    {
      _i += len;
      return true;
    }
    return false;
  }

  static void append_utf8(std::string& out, unsigned cp)
  {
    if (cp < 0x80)
      out.push_back((char)cp);
    else if (cp < 0x800)
    {
      out.push_back((char)(0xC0 | (cp >> 6)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    }
    else if (cp < 0x10000)
    {
      out.push_back((char)(0xE0 | (cp >> 12)));
      out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    }
    else
    {
      out.push_back((char)(0xF0 | (cp >> 18)));
      out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    }
  }

  unsigned parse_hex4()
  {
    if (_n - _i < 4)
      fail("truncated \\u escape");
    unsigned v = 0;
    for (int k = 0; k < 4; k++)
    {
      const char c = _s[_i++];
      v <<= 4;
      if (c >= '0' && c <= '9')
        v |= (unsigned)(c - '0');
      else if (c >= 'a' && c <= 'f')
        v |= (unsigned)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F')
        v |= (unsigned)(c - 'A' + 10);
      else
        fail("bad hex digit in \\u escape");
    }
    return v;
  }

  std::string parse_string_body()
  {
    // precondition: opening quote consumed
    std::string out;
    while (true)
    {
      if (_i >= _n)
        fail("unterminated string");
      const char c = _s[_i++];
      if (c == '"')
        return out;
      if (c != '\\')
      {
        out.push_back(c);
        continue;
      }
      if (_i >= _n)
        fail("unterminated escape");
      const char e = _s[_i++];
      switch (e)
      {
        case '"': out.push_back('"'); break;
        case '\\': out.push_back('\\'); break;
        case '/': out.push_back('/'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'n': out.push_back('\n'); break;
        case 'r': out.push_back('\r'); break;
        case 't': out.push_back('\t'); break;
        case 'u':
        {
          unsigned cp = parse_hex4();
          if (cp >= 0xD800 && cp <= 0xDBFF && _n - _i >= 6 && _s[_i] == '\\' && _s[_i + 1] == 'u')
          {
            _i += 2;
            const unsigned lo = parse_hex4();
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          append_utf8(out, cp);
          break;
        }
        default: fail("bad escape character");
      }
    }
  }

  Value parse_number()
  {
    const size_t start = _i;
    if (_i < _n && (_s[_i] == '-' || _s[_i] == '+'))
      _i++;
    // tolerate NaN / Infinity as python's json.dump may emit them
    if (consume("NaN"))
    {
      Value v;
      v._type = Value::Type::Number;
      v._number = std::nan("");
      return v;
    }
    if (consume("Infinity"))
    {
      Value v;
      v._type = Value::Type::Number;
      v._number = (_s[start] == '-') ? -INFINITY : INFINITY;
      return v;
    }
    while (_i < _n
           && ((_s[_i] >= '0' && _s[_i] <= '9') || _s[_i] == '.' || _s[_i] == 'e' || _s[_i] == 'E' || _s[_i] == '+'
               || _s[_i] == '-'))
      _i++;
    if (_i == start)
      fail("expected a value");
    const std::string tok(_s + start, _i - start);
    // std::from_chars: locale-independent (strtod reads "0,5" in a decimal-comma locale, and a DAW host may well have
    // called setlocale); nlohmann::json, the reference's parser, is locale-independent too
    double d = 0.0;
    const char* first = tok.c_str() + ((tok[0] == '+') ? 1 : 0);
    const auto res = std::from_chars(first, tok.c_str() + tok.size(), d);
    if (res.ec == std::errc::result_out_of_range && res.ptr == tok.c_str() + tok.size())
    {
      // beyond double's range: underflow to zero / overflow to infinity, like strtod and nlohmann::json
      const size_t e = tok.find_first_of("eE");
      const bool tiny = e != std::string::npos && e + 1 < tok.size() && tok[e + 1] == '-';
      d = tiny ? 0.0 : INFINITY;
      if (tok[0] == '-')
        d = -d;
    }
    else if (res.ec != std::errc() || res.ptr != tok.c_str() + tok.size())
      fail("malformed number '" + tok + "'");
    Value v;
    v._type = Value::Type::Number;
    v._number = d;
    return v;
  }

  Value parse_value(int depth)
  {
    if (depth > 256)
      fail("nesting too deep");
    skip_ws();
    if (_i >= _n)
      fail("unexpected end of input");
    const char c = _s[_i];
    Value v;
    if (c == '{')
    {
      _i++;
      v._type = Value::Type::Object;
      skip_ws();
      if (_i < _n && _s[_i] == '}')
      {
        _i++;
        return v;
      }
      while (true)
      {
        skip_ws();
        if (_i >= _n || _s[_i] != '"')
          fail("expected string key");
        _i++;
        std::string key = parse_string_body();
        skip_ws();
        if (_i >= _n || _s[_i] != ':')
          fail("expected ':'");
        _i++;
        Value child = parse_value(depth + 1);
        // duplicate keys: last one wins, like nlohmann
        bool replaced = false;
        for (auto& kv : v._members)
          if (kv.first == key)
          {
            kv.second = std::move(child);
            replaced = true;
            break;
          }
        if (!replaced)
          v._members.emplace_back(std::move(key), std::move(child));
        skip_ws();
        if (_i < _n && _s[_i] == ',')
        {
          _i++;
          continue;
        }
        if (_i < _n && _s[_i] == '}')
        {
          _i++;
          return v;
        }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[')
    {
      _i++;
      v._type = Value::Type::Array;
      skip_ws();
      if (_i < _n && _s[_i] == ']')
      {
        _i++;
        return v;
      }
      while (true)
      {
        v._items.push_back(parse_value(depth + 1));
        skip_ws();
        if (_i < _n && _s[_i] == ',')
        {
          _i++;
          continue;
        }
        if (_i < _n && _s[_i] == ']')
        {
          _i++;
          return v;
        }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"')
    {
      _i++;
      v._type = Value::Type::String;
      v._string = parse_string_body();
      return v;
    }
    if (consume("true"))
    {
      v._type = Value::Type::Bool;
      v._bool = true;
      return v;
    }
    if (consume("false"))
    {
      v._type = Value::Type::Bool;
      v._bool = false;
      return v;
    }
    if (consume("null"))
      return v;
    return parse_number();
  }
};

Value Value::parse(const std::string& text)
{
  Parser p(text);
  return p.parse_document();
}

namespace
{
void dump_string(const std::string& s, std::string& out)
{
  out.push_back('"');
  for (const char ch : s)
  {
    switch (ch)
    {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      default:
        if ((unsigned char)ch < 0x20)
        {
          char buf[8];
          std::snprintf(buf, sizeof(buf), "\\u%04x", (unsigned)(unsigned char)ch);
          out += buf;
        }
        else
          out.push_back(ch);
    }
  }
  out.push_back('"');
}

void dump_value(const Value& v, std::string& out)
{
  switch (v.type())
  {
    case Value::Type::Null: out += "null"; break;
    case Value::Type::Bool: out += v.as_bool() ? "true" : "false"; break;
    case Value::Type::Number:
    {
      out += number_to_string(v.as_double());
      break;
    }
    case Value::Type::String: dump_string(v.as_string(), out); break;
    case Value::Type::Array:
    {
      out.push_back('[');
      bool first = true;
      for (const auto& it : v.items())
      {
        if (!first)
          out.push_back(',');
        first = false;
        dump_value(it, out);
      }
      out.push_back(']');
      break;
    }
    case Value::Type::Object:
    {
      out.push_back('{');
      bool first = true;
      for (const auto& kv : v.members())
      {
        if (!first)
          out.push_back(',');
        first = false;
        dump_string(kv.first, out);
        out.push_back(':');
        dump_value(kv.second, out);
      }
      out.push_back('}');
      break;
    }
  }
}
} // namespace

std::string Value::dump() const
{
  std::string out;
  dump_value(*this, out);
  return out;
}

bool Value::as_bool(const char* what) const
{
  if (_type == Type::Bool)
    return _bool;
  if (_type == Type::Number)
    return _number != 0.0;
  throw std::runtime_error(std::string("JSON: expected a boolean for ") + what);
}

double Value::as_double(const char* what) const
{
  if (_type == Type::Number)
    return _number;
  if (_type == Type::Bool)
    return _bool ? 1.0 : 0.0;
  throw std::runtime_error(std::string("JSON: expected a number for ") + what);
}

int Value::as_int(const char* what) const
{
  return (int)as_double(what);
}

const std::string& Value::as_string(const char* what) const
{
  if (_type != Type::String)
    throw std::runtime_error(std::string("JSON: expected a string for ") + what);
  return _string;
}

const std::vector<Value>& Value::items(const char* what) const
{
  if (_type != Type::Array)
    throw std::runtime_error(std::string("JSON: expected an array for ") + what);
  return _items;
}

bool Value::contains(const std::string& key) const
{
  if (_type != Type::Object)
    return false;
  for (const auto& kv : _members)
    if (kv.first == key)
      return true;
  return false;
}

const Value& Value::get(const std::string& key) const
{
  if (_type == Type::Object)
    for (const auto& kv : _members)
      if (kv.first == key)
        return kv.second;
  return kNull;
}

const Value& Value::at(const std::string& key) const
{
  if (_type == Type::Object)
    for (const auto& kv : _members)
      if (kv.first == key)
        return kv.second;
  throw std::runtime_error("JSON: missing key \"" + key + "\"");
}

int Value::value_int(const std::string& key, int dflt) const
{
  return contains(key) ? at(key).as_int(key.c_str()) : dflt;
}

double Value::value_double(const std::string& key, double dflt) const
{
  return contains(key) ? at(key).as_double(key.c_str()) : dflt;
}

bool Value::value_bool(const std::string& key, bool dflt) const
{
  return contains(key) ? at(key).as_bool(key.c_str()) : dflt;
}

} // namespace json
} // namespace namb200

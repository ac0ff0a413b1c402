// json_mini.h -- a small self-contained JSON reader/writer for the configuration documents the reference accepts
// (nlohmann::json in the reference; the C ABI takes the same documents as text). Supports the full JSON grammar plus
// the // and /* */ comments the reference's loader tolerates (json::parse(f, nullptr, true, true)).
#pragma once
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace tcnnb {
namespace json {

struct Value {
	enum Type { Null, Bool, Number, String, Array, Object };
	Type type = Null;
	bool b = false;
	double num = 0.0;
	std::string str;
	std::vector<Value> arr;
	std::vector<std::pair<std::string, Value>> obj;  // insertion order preserved

	Value() {}
	static Value object() { Value v; v.type = Object; return v; }
	static Value number(double d) { Value v; v.type = Number; v.num = d; return v; }
	static Value string(const std::string& s) { Value v; v.type = String; v.str = s; return v; }
	static Value boolean(bool x) { Value v; v.type = Bool; v.b = x; return v; }

	bool is_object() const { return type == Object; }
	bool contains(const std::string& key) const { return find(key) != nullptr; }
	const Value* find(const std::string& key) const {
		if (type != Object) return nullptr;
		for (auto& kv : obj) if (kv.first == key) return &kv.second;
		return nullptr;
	}
	Value& operator[](const std::string& key) {
		if (type == Null) type = Object;
		for (auto& kv : obj) if (kv.first == key) return kv.second;
		obj.emplace_back(key, Value{});
		return obj.back().second;
	}
	// .value(key, default) lookups with the reference's semantics (missing key -> default).
	const Value& sub(const std::string& key) const {
		static const Value empty = Value::object();
		const Value* v = find(key);
		return v ? *v : empty;
	}
	double value(const std::string& key, double def) const {
		const Value* v = find(key);
		if (!v) return def;
		if (v->type == Number) return v->num;
		if (v->type == Bool) return v->b ? 1.0 : 0.0;
		throw std::runtime_error("JSON: key '" + key + "' is not a number");
	}
	bool value(const std::string& key, bool def) const {
		const Value* v = find(key);
		if (!v) return def;
		if (v->type == Bool) return v->b;
		if (v->type == Number) return v->num != 0.0;
		throw std::runtime_error("JSON: key '" + key + "' is not a boolean");
	}
	std::string value(const std::string& key, const std::string& def) const {
		const Value* v = find(key);
		if (!v) return def;
		if (v->type == String) return v->str;
		throw std::runtime_error("JSON: key '" + key + "' is not a string");
	}
	std::string value(const std::string& key, const char* def) const { return value(key, std::string(def)); }
};

namespace detail {

struct Parser {
	const std::string& s;
	size_t i = 0;
	explicit Parser(const std::string& text) : s(text) {}

	[[noreturn]] void fail(const std::string& what) const {
		throw std::runtime_error("JSON parse error at offset " + std::to_string(i) + ": " + what);
	}
	void skip_ws() {
		for (;;) {
			while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
			if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/') {
				while (i < s.size() && s[i] != '\n') ++i;
			} else if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '*') {
				i += 2;
				while (i + 1 < s.size() && !(s[i] == '*' && s[i + 1] == '/')) ++i;
				i += 2;
			} else {
				return;
			}
		}
	}
	Value parse_value() {
		skip_ws();
		if (i >= s.size()) fail("unexpected end of input");
		const char c = s[i];
		if (c == '{') return parse_object();
		if (c == '[') return parse_array();
		if (c == '"') return Value::string(parse_string());
		if (s.compare(i, 4, "true") == 0) { i += 4; return Value::boolean(true); }
		if (s.compare(i, 5, "false") == 0) { i += 5; return Value::boolean(false); }
		if (s.compare(i, 4, "null") == 0) { i += 4; return Value{}; }
		return parse_number();
	}
	Value parse_number() {
		const char* begin = s.c_str() + i;
		char* end = nullptr;
		const double d = std::strtod(begin, &end);
		if (end == begin) fail("invalid value");
		i += (size_t)(end - begin);
		return Value::number(d);
	}
	std::string parse_string() {
		++i;  // opening quote
		std::string out;
		while (i < s.size() && s[i] != '"') {
			char c = s[i++];
			if (c == '\\') {
				if (i >= s.size()) fail("bad escape");
				const char e = s[i++];
				switch (e) {
					case 'n': out += '\n'; break;
					case 't': out += '\t'; break;
					case 'r': out += '\r'; break;
					case 'b': out += '\b'; break;
					case 'f': out += '\f'; break;
					case 'u': {
						if (i + 4 > s.size()) fail("bad \\u escape");
						const unsigned cp = (unsigned)std::strtoul(s.substr(i, 4).c_str(), nullptr, 16);
						i += 4;
						if (cp < 0x80) out += (char)cp;
						else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
						else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
						break;
					}
					default: out += e;
				}
			} else {
				out += c;
			}
		}
		if (i >= s.size()) fail("unterminated string");
		++i;
		return out;
	}
	Value parse_array() {
		Value v;
		v.type = Value::Array;
		++i;
		skip_ws();
		if (i < s.size() && s[i] == ']') { ++i; return v; }
		for (;;) {
			v.arr.push_back(parse_value());
			skip_ws();
			if (i < s.size() && s[i] == ',') { ++i; continue; }
			if (i < s.size() && s[i] == ']') { ++i; return v; }
			fail("expected ',' or ']'");
		}
	}
	Value parse_object() {
		Value v = Value::object();
		++i;
		skip_ws();
		if (i < s.size() && s[i] == '}') { ++i; return v; }
		for (;;) {
			skip_ws();
			if (i >= s.size() || s[i] != '"') fail("expected string key");
			std::string key = parse_string();
			skip_ws();
			if (i >= s.size() || s[i] != ':') fail("expected ':'");
			++i;
			v[key] = parse_value();
			skip_ws();
			if (i < s.size() && s[i] == ',') { ++i; continue; }
			if (i < s.size() && s[i] == '}') { ++i; return v; }
			fail("expected ',' or '}'");
		}
	}
};

inline void dump(const Value& v, std::ostringstream& o) {
	switch (v.type) {
		case Value::Null: o << "null"; break;
		case Value::Bool: o << (v.b ? "true" : "false"); break;
		case Value::Number: {
			if (std::floor(v.num) == v.num && std::fabs(v.num) < 1e15) {
				o << (long long)v.num;
			} else {
				char buf[40];
				std::snprintf(buf, sizeof(buf), "%.9g", v.num);
				o << buf;
			}
			break;
		}
		case Value::String: {
			o << '"';
			for (char c : v.str) {
				if (c == '"' || c == '\\') o << '\\' << c;
				else if (c == '\n') o << "\\n";
				else o << c;
			}
			o << '"';
			break;
		}
		case Value::Array: {
			o << '[';
			for (size_t k = 0; k < v.arr.size(); ++k) { if (k) o << ", "; dump(v.arr[k], o); }
			o << ']';
			break;
		}
		case Value::Object: {
			o << '{';
			for (size_t k = 0; k < v.obj.size(); ++k) {
				if (k) o << ", ";
				o << '"' << v.obj[k].first << "\": ";
				dump(v.obj[k].second, o);
			}
			o << '}';
			break;
		}
	}
}

}  // namespace detail

inline Value parse(const std::string& text) {
	detail::Parser p(text);
	Value v = p.parse_value();
	p.skip_ws();
	if (p.i != text.size()) p.fail("trailing characters");
	return v;
}

inline std::string dump(const Value& v) {
	std::ostringstream o;
	detail::dump(v, o);
	return o.str();
}

}  // namespace json
}  // namespace tcnnb

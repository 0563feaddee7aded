// Minimal JSON DOM for model files (.nam / keras .json / .aidax).  The reference parses these with
// nlohmann::json (NeuralModel.cpp:330-336); this reader keeps the behaviours the model path depends on:
// std::map-like sorted object keys (metadata iteration order, NeuralModelImpl.h:85-94), integer vs float
// distinction, and a compact dump() that matches nlohmann's for the values GetMetadata returns.
#pragma once
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace nab200
{
	class Json
	{
	public:
		enum Type { Null, Bool, Int, Float, String, Array, Object };

		Type type = Null;
		bool b = false;
		int64_t i = 0;
		double d = 0.0;
		std::string s;
		std::vector<Json> arr;
		std::map<std::string, Json> obj;   // sorted keys, like nlohmann's default object_t

		bool is_null() const { return type == Null; }
		bool is_number() const { return type == Int || type == Float; }
		bool is_array() const { return type == Array; }
		bool is_object() const { return type == Object; }
		bool is_string() const { return type == String; }
		bool is_bool() const { return type == Bool; }

		bool contains(const std::string& key) const { return type == Object && obj.find(key) != obj.end(); }

		const Json& at(const std::string& key) const
		{
			if (type != Object) throw std::runtime_error("json: not an object (key '" + key + "')");
			auto it = obj.find(key);
			if (it == obj.end()) throw std::out_of_range("json: key '" + key + "' not found");
			return it->second;
		}

		const Json& at(size_t idx) const
		{
			if (type != Array) throw std::runtime_error("json: not an array");
			if (idx >= arr.size()) throw std::out_of_range("json: array index out of range");
			return arr[idx];
		}

		size_t size() const { return type == Array ? arr.size() : (type == Object ? obj.size() : (type == Null ? 0 : 1)); }

		double as_double() const
		{
			if (type == Int) return (double)i;
			if (type == Float) return d;
			if (type == Bool) return b ? 1.0 : 0.0;
			throw std::runtime_error("json: not a number");
		}

		float as_float() const { return (float)as_double(); }

		int as_int() const
		{
			if (type == Int) return (int)i;
			if (type == Float) return (int)d;
			if (type == Bool) return b ? 1 : 0;
			throw std::runtime_error("json: not a number");
		}

		bool as_bool() const
		{
			if (type == Bool) return b;
			if (type == Int) return i != 0;
			if (type == Float) return d != 0.0;
			throw std::runtime_error("json: not a boolean");
		}

		const std::string& as_string() const
		{
			if (type != String) throw std::runtime_error("json: not a string");
			return s;
		}

		int value_int(const std::string& key, int def) const
		{
			if (!contains(key)) return def;
			const Json& v = at(key);
			return v.is_number() ? v.as_int() : def;
		}

		bool value_bool(const std::string& key, bool def) const
		{
			if (!contains(key)) return def;
			const Json& v = at(key);
			return (v.is_bool() || v.is_number()) ? v.as_bool() : def;
		}

		double value_double(const std::string& key, double def) const
		{
			if (!contains(key)) return def;
			const Json& v = at(key);
			return v.is_number() ? v.as_double() : def;
		}

		// compact serialisation, nlohmann::json::dump() conventions
		std::string dump() const
		{
			std::string out;
			dump_to(out);
			return out;
		}

		static Json parse(const std::string& text)
		{
			Parser p{ text.data(), text.data() + text.size() };
			Json v = p.value();
			p.ws();
			if (p.cur != p.end) throw std::runtime_error("json: trailing characters");
			return v;
		}

	private:
		static void dump_string(const std::string& str, std::string& out)
		{
			out.push_back('"');
			for (unsigned char c : str)
			{
				switch (c)
				{
				case '"': out += "\\\""; break;
				case '\\': out += "\\\\"; break;
				case '\b': out += "\\b"; break;
				case '\f': out += "\\f"; break;
				case '\n': out += "\\n"; break;
				case '\r': out += "\\r"; break;
				case '\t': out += "\\t"; break;
				default:
					if (c < 0x20)
					{
						char buf[8];
						snprintf(buf, sizeof(buf), "\\u%04x", c);
						out += buf;
					}
					else out.push_back((char)c);
				}
			}
			out.push_back('"');
		}

		void dump_to(std::string& out) const
		{
			switch (type)
			{
			case Null: out += "null"; break;
			case Bool: out += b ? "true" : "false"; break;
			case Int: out += std::to_string(i); break;
			case Float:
			{
				if (!std::isfinite(d)) { out += "null"; break; }
				char buf[64];
				auto res = std::to_chars(buf, buf + sizeof(buf), d);   // shortest round-trip, like nlohmann's Grisu2
				std::string t(buf, res.ptr);
				if (t.find_first_of(".eE") == std::string::npos) t += ".0";
				out += t;
				break;
			}
			case String: dump_string(s, out); break;
			case Array:
				out.push_back('[');
				for (size_t k = 0; k < arr.size(); k++)
				{
					if (k) out.push_back(',');
					arr[k].dump_to(out);
				}
				out.push_back(']');
				break;
			case Object:
			{
				out.push_back('{');
				bool first = true;
				for (auto& kv : obj)
				{
					if (!first) out.push_back(',');
					first = false;
					dump_string(kv.first, out);
					out.push_back(':');
					kv.second.dump_to(out);
				}
				out.push_back('}');
				break;
			}
			}
		}

		struct Parser
		{
			const char* cur;
			const char* end;

			void ws()
			{
				while (cur < end && (*cur == ' ' || *cur == '\t' || *cur == '\n' || *cur == '\r')) cur++;
			}

			[[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("json: parse error: ") + what); }

			Json value()
			{
				ws();
				if (cur >= end) fail("unexpected end");
				switch (*cur)
				{
				case '{': return object();
				case '[': return array();
				case '"': { Json v; v.type = String; v.s = string(); return v; }
				case 't': lit("true"); { Json v; v.type = Bool; v.b = true; return v; }
				case 'f': lit("false"); { Json v; v.type = Bool; v.b = false; return v; }
				case 'n': lit("null"); return Json();
				case 'N': lit("NaN"); { Json v; v.type = Float; v.d = NAN; return v; }   // Python json.dump can emit these
				case 'I': lit("Infinity"); { Json v; v.type = Float; v.d = INFINITY; return v; }
				default: return number();
				}
			}

			void lit(const char* w)
			{
				size_t n = strlen(w);
				if ((size_t)(end - cur) < n || strncmp(cur, w, n) != 0) fail("bad literal");
				cur += n;
			}

			Json number()
			{
				const char* start = cur;
				if (cur < end && *cur == '-')
				{
					cur++;
					if (cur < end && *cur == 'I') { lit("Infinity"); Json v; v.type = Float; v.d = -INFINITY; return v; }
				}
				bool isFloat = false;
				while (cur < end && ((*cur >= '0' && *cur <= '9') || *cur == '.' || *cur == 'e' || *cur == 'E' || *cur == '+' || *cur == '-'))
				{
					if (*cur == '.' || *cur == 'e' || *cur == 'E') isFloat = true;
					cur++;
				}
				if (cur == start) fail("unexpected character");
				Json v;
				if (!isFloat)
				{
					int64_t iv = 0;
					auto res = std::from_chars(start, cur, iv);
					if (res.ec == std::errc() && res.ptr == cur) { v.type = Int; v.i = iv; return v; }
				}
				double dv = 0.0;
				auto res = std::from_chars(start, cur, dv);
				if (res.ec != std::errc() || res.ptr != cur) fail("bad number");
				v.type = Float;
				v.d = dv;
				return v;
			}

			static void utf8(uint32_t cp, std::string& out)
			{
				if (cp < 0x80) out.push_back((char)cp);
				else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
				else if (cp < 0x10000)
				{
					out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
				}
				else
				{
					out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
					out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
				}
			}

			uint32_t hex4()
			{
				if (end - cur < 4) fail("bad \\u escape");
				uint32_t v = 0;
				for (int k = 0; k < 4; k++)
				{
					char c = *cur++;
					v <<= 4;
					if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
					else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
					else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
					else fail("bad \\u escape");
				}
				return v;
			}

			std::string string()
			{
				std::string out;
				cur++;   // opening quote
				while (true)
				{
					if (cur >= end) fail("unterminated string");
					char c = *cur++;
					if (c == '"') break;
					if (c != '\\') { out.push_back(c); continue; }
					if (cur >= end) fail("bad escape");
					char e = *cur++;
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
						uint32_t cp = hex4();
						if (cp >= 0xD800 && cp <= 0xDBFF && end - cur >= 6 && cur[0] == '\\' && cur[1] == 'u')
						{
							cur += 2;
							uint32_t lo = hex4();
							cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
						}
						utf8(cp, out);
						break;
					}
					default: fail("bad escape");
					}
				}
				return out;
			}

			Json array()
			{
				Json v;
				v.type = Array;
				cur++;
				ws();
				if (cur < end && *cur == ']') { cur++; return v; }
				while (true)
				{
					v.arr.push_back(value());
					ws();
					if (cur >= end) fail("unterminated array");
					if (*cur == ',') { cur++; continue; }
					if (*cur == ']') { cur++; break; }
					fail("expected , or ]");
				}
				return v;
			}

			Json object()
			{
				Json v;
				v.type = Object;
				cur++;
				ws();
				if (cur < end && *cur == '}') { cur++; return v; }
				while (true)
				{
					ws();
					if (cur >= end || *cur != '"') fail("expected string key");
					std::string key = string();
					ws();
					if (cur >= end || *cur != ':') fail("expected :");
					cur++;
					v.obj[key] = value();
					ws();
					if (cur >= end) fail("unterminated object");
					if (*cur == ',') { cur++; continue; }
					if (*cur == '}') { cur++; break; }
					fail("expected , or }");
				}
				return v;
			}
		};
	};
}

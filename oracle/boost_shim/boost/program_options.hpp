// Test-infrastructure only: small stand-in for <boost/program_options.hpp>, just enough for the
// reference's `prep` and `junc` command lines (oracle build, see oracle/README.md).
#pragma once
#include <boost/lexical_cast.hpp>   // real Boost.Program_options pulls it in; src/bam_filter.cc relies on that
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <sstream>
#include <iostream>
#include <stdexcept>
#include <typeinfo>
#include <type_traits>
namespace boost { namespace program_options {

class error : public std::logic_error { public: error(const std::string& w) : std::logic_error(w) {} };

namespace shim_detail {
template <class T> struct conv {
    static T from(const std::string& s) {
        if constexpr (std::is_constructible<T, std::string>::value) { return T(s); }
        else {
            std::istringstream ss(s); T t;
            if constexpr (sizeof(T) == 1) { int x; if (!(ss >> x)) throw error("bad value: " + s); t = (T)x; }
            else { if (!(ss >> t)) throw error("the argument ('" + s + "') is invalid"); }
            return t;
        }
    }
};
template <class T> struct is_vector : std::false_type {};
template <class T, class A> struct is_vector<std::vector<T, A>> : std::true_type {};
}

class value_semantic {
public:
    virtual ~value_semantic() {}
    virtual bool is_switch() const = 0;
    virtual bool multi() const = 0;
    virtual void parse(const std::string& s) = 0;   // one token
    virtual void apply_default() = 0;
    virtual bool has_default() const = 0;
    virtual std::string default_text() const = 0;
    virtual const void* ptr() const = 0;
    virtual const std::type_info& type() const = 0;
};

template <class T> class typed_value : public value_semantic {
public:
    explicit typed_value(T* store) : store_(store) {}
    typed_value* default_value(const T& v) {
        def_ = std::make_shared<T>(v);
        std::ostringstream ss;
        if constexpr (!shim_detail::is_vector<T>::value) { if constexpr (sizeof(T) == 1 && std::is_integral<T>::value) ss << (int)v; else ss << v; }
        deftxt_ = ss.str();
        return this;
    }
    typed_value* default_value(const T& v, const std::string& txt) { def_ = std::make_shared<T>(v); deftxt_ = txt; return this; }
    typed_value* required() { return this; }
    typed_value* multitoken() { return this; }
    typed_value* implicit_value(const T&) { return this; }
    bool is_switch() const override { return switch_; }
    bool multi() const override { return shim_detail::is_vector<T>::value; }
    void parse(const std::string& s) override {
        if constexpr (shim_detail::is_vector<T>::value) {
            held_.push_back(shim_detail::conv<typename T::value_type>::from(s));
        } else if constexpr (std::is_same<T, bool>::value) {
            held_ = switch_ ? true : (s == "1" || s == "true" || s == "on" || s == "yes");
        } else {
            held_ = shim_detail::conv<T>::from(s);
        }
        if (store_) *store_ = held_;
    }
    void apply_default() override { if (def_) { held_ = *def_; if (store_) *store_ = held_; } }
    bool has_default() const override { return (bool)def_; }
    std::string default_text() const override { return deftxt_; }
    const void* ptr() const override { return &held_; }
    const std::type_info& type() const override { return typeid(T); }
    bool switch_ = false;
private:
    T* store_; T held_{}; std::shared_ptr<T> def_; std::string deftxt_;
};

template <class T> inline typed_value<T>* value() { return new typed_value<T>(nullptr); }
template <class T> inline typed_value<T>* value(T* v) { return new typed_value<T>(v); }
inline typed_value<bool>* bool_switch(bool* v = nullptr) {
    typed_value<bool>* t = new typed_value<bool>(v); t->switch_ = true; t->default_value(false); return t;
}

struct option_description {
    std::string long_name, short_name, description;
    std::shared_ptr<value_semantic> semantic;
};

class options_description;
class options_description_easy_init {
public:
    explicit options_description_easy_init(options_description* o) : owner_(o) {}
    options_description_easy_init& operator()(const char* name, const char* desc);
    options_description_easy_init& operator()(const char* name, value_semantic* s, const char* desc = "");
private:
    options_description* owner_;
};

class options_description {
public:
    options_description() {}
    explicit options_description(const std::string& caption, unsigned = 80, unsigned = 40) : caption_(caption) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    options_description& add(const options_description& o) {
        groups_.push_back(o);
        for (const auto& d : o.all_) all_.push_back(d);
        return *this;
    }
    void add_option(const std::shared_ptr<option_description>& d) { own_.push_back(d); all_.push_back(d); }
    const std::vector<std::shared_ptr<option_description>>& options() const { return all_; }
    std::shared_ptr<option_description> find_long(const std::string& n) const {
        for (const auto& d : all_) if (d->long_name == n) return d; return nullptr;
    }
    std::shared_ptr<option_description> find_short(const std::string& n) const {
        for (const auto& d : all_) if (!d->short_name.empty() && d->short_name == n) return d; return nullptr;
    }
    void print(std::ostream& os) const {
        if (!caption_.empty()) os << caption_ << ":\n";
        for (const auto& d : own_) {
            std::string left = "  ";
            if (!d->short_name.empty()) left += "-" + d->short_name + " [ --" + d->long_name + " ]";
            else left += "--" + d->long_name;
            if (d->semantic && !d->semantic->is_switch()) {
                left += " arg";
                if (d->semantic->has_default()) left += " (=" + d->semantic->default_text() + ")";
            }
            os << left << "\n        " << d->description << "\n";
        }
        for (const auto& g : groups_) { g.print(os); os << "\n"; }
    }
private:
    std::string caption_;
    std::vector<std::shared_ptr<option_description>> own_, all_;
    std::vector<options_description> groups_;
};
inline std::ostream& operator<<(std::ostream& os, const options_description& d) { d.print(os); return os; }

inline options_description_easy_init& options_description_easy_init::operator()(const char* name, const char* desc) {
    return (*this)(name, bool_switch(), desc);
}
inline options_description_easy_init& options_description_easy_init::operator()(const char* name, value_semantic* s, const char* desc) {
    auto d = std::make_shared<option_description>();
    std::string n(name); size_t c = n.find(',');
    if (c == std::string::npos) d->long_name = n; else { d->long_name = n.substr(0, c); d->short_name = n.substr(c + 1); }
    d->description = desc ? desc : ""; d->semantic.reset(s);
    owner_->add_option(d);
    return *this;
}

class positional_options_description {
public:
    positional_options_description& add(const char* name, int max_count) { names_.push_back({name, max_count}); return *this; }
    std::string name_for_position(size_t pos) const {
        size_t acc = 0;
        for (const auto& n : names_) { if (n.second < 0) return n.first; acc += (size_t)n.second; if (pos < acc) return n.first; }
        return names_.empty() ? std::string() : names_.back().first;
    }
    bool empty() const { return names_.empty(); }
private:
    std::vector<std::pair<std::string, int>> names_;
};

struct parsed_options {
    const options_description* desc = nullptr;
    std::vector<std::pair<std::shared_ptr<option_description>, std::string>> hits;  // option, token ("" for switches)
};

class command_line_parser {
public:
    command_line_parser(int argc, const char* const* argv) { for (int i = 1; i < argc; i++) args_.push_back(argv[i]); }
    command_line_parser(int argc, char** argv) { for (int i = 1; i < argc; i++) args_.push_back(argv[i]); }
    command_line_parser& options(const options_description& d) { desc_ = &d; return *this; }
    command_line_parser& positional(const positional_options_description& p) { pos_ = &p; return *this; }
    command_line_parser& allow_unregistered() { allow_unreg_ = true; return *this; }
    parsed_options run() {
        parsed_options out; out.desc = desc_;
        size_t npos = 0; bool only_pos = false;
        for (size_t i = 0; i < args_.size(); i++) {
            const std::string& a = args_[i];
            std::shared_ptr<option_description> d; std::string val; bool have_val = false;
            if (!only_pos && a == "--") { only_pos = true; continue; }
            if (!only_pos && a.size() > 2 && a[0] == '-' && a[1] == '-') {
                std::string n = a.substr(2); size_t eq = n.find('=');
                if (eq != std::string::npos) { val = n.substr(eq + 1); n = n.substr(0, eq); have_val = true; }
                d = desc_->find_long(n);
                if (!d) { if (allow_unreg_) continue; throw error("unrecognised option '" + a + "'"); }
            } else if (!only_pos && a.size() >= 2 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
                d = desc_->find_short(a.substr(1, 1));
                if (!d) { if (allow_unreg_) continue; throw error("unrecognised option '" + a + "'"); }
                if (a.size() > 2) { val = a.substr(2); have_val = true; }
            } else {
                if (!pos_ || pos_->empty()) { if (allow_unreg_) continue; throw error("too many positional options have been specified on the command line"); }
                d = desc_->find_long(pos_->name_for_position(npos++));
                if (!d) throw error("unknown positional option");
                out.hits.push_back({d, a});
                continue;
            }
            if (d->semantic->is_switch()) { out.hits.push_back({d, ""}); continue; }
            if (!have_val) {
                if (i + 1 >= args_.size()) throw error("the required argument for option '--" + d->long_name + "' is missing");
                val = args_[++i];
            }
            out.hits.push_back({d, val});
        }
        return out;
    }
private:
    std::vector<std::string> args_;
    const options_description* desc_ = nullptr;
    const positional_options_description* pos_ = nullptr;
    bool allow_unreg_ = false;
};
inline parsed_options parse_command_line(int argc, char** argv, const options_description& d) {
    return command_line_parser(argc, argv).options(d).run();
}

class variable_value {
public:
    variable_value() {}
    explicit variable_value(std::shared_ptr<value_semantic> s) : s_(s) {}
    template <class T> const T& as() const {
        if (!s_ || s_->type() != typeid(T)) throw error("bad any_cast in variables_map");
        return *static_cast<const T*>(s_->ptr());
    }
    bool empty() const { return !s_; }
private:
    std::shared_ptr<value_semantic> s_;
};

class variables_map {
public:
    size_t count(const std::string& n) const { return m_.count(n); }
    const variable_value& operator[](const std::string& n) const {
        static variable_value none; auto it = m_.find(n); return it == m_.end() ? none : it->second;
    }
    std::map<std::string, variable_value> m_;
};

inline void store(const parsed_options& p, variables_map& vm) {
    for (const auto& h : p.hits) {
        h.first->semantic->parse(h.second);
        vm.m_[h.first->long_name] = variable_value(h.first->semantic);
    }
    if (p.desc) for (const auto& d : p.desc->options()) {
        if (!vm.m_.count(d->long_name) && d->semantic->has_default()) {
            d->semantic->apply_default();
            vm.m_[d->long_name] = variable_value(d->semantic);
        }
    }
}
inline void notify(variables_map&) {}
}}

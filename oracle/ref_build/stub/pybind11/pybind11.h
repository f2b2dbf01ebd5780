// TEST INFRASTRUCTURE (oracle/_ref build only): permissive stand-in so that the reference's headers,
// which include pybind11 for their Export_* registration functions, compile without Python.  The
// registration functions are parsed but never called; nothing here does anything.
#pragma once
#include <Eigen/Core>
#include <string>
#include <vector>
#include <map>
namespace pybind11 {
struct module {
    template <class... A> module& def(A&&...) { return *this; }
    template <class... A> module def_submodule(A&&...) { return *this; }
    template <class... A> module& attr(A&&...) { return *this; }
};
using module_ = module;
template <class... T> struct init { };
struct is_operator {};
struct arg { arg(const char*) {} template <class T> arg& operator=(T&&) { return *this; } };
struct self_t {};
static const self_t self{};
struct op_ {};
#define CIPC_STUB_OP(o) \
    inline op_ operator o(const self_t&, const self_t&) { return op_{}; } \
    template <class T> op_ operator o(const self_t&, const T&) { return op_{}; } \
    template <class T> op_ operator o(const T&, const self_t&) { return op_{}; }
CIPC_STUB_OP(+) CIPC_STUB_OP(-) CIPC_STUB_OP(*) CIPC_STUB_OP(/)
#undef CIPC_STUB_OP
#define CIPC_STUB_IOP(o) \
    inline op_ operator o(const self_t&, const self_t&) { return op_{}; } \
    template <class T> op_ operator o(const self_t&, const T&) { return op_{}; }
CIPC_STUB_IOP(+=) CIPC_STUB_IOP(-=) CIPC_STUB_IOP(*=) CIPC_STUB_IOP(/=)
#undef CIPC_STUB_IOP
inline op_ operator-(const self_t&) { return op_{}; }
template <class... T> struct class_ {
    template <class... A> class_(A&&...) {}
    template <class... A> class_& def(A&&...) { return *this; }
    template <class... A> class_& def_readwrite(A&&...) { return *this; }
    template <class... A> class_& def_readonly(A&&...) { return *this; }
    template <class... A> class_& def_static(A&&...) { return *this; }
    template <class... A> class_& def_property(A&&...) { return *this; }
    template <class... A> class_& def_property_readonly(A&&...) { return *this; }
};
template <class V, class... A> class_<V> bind_vector(A&&...) { return class_<V>(); }
template <class V, class... A> class_<V> bind_map(A&&...) { return class_<V>(); }
struct scoped_ostream_redirect { template <class... A> scoped_ostream_redirect(A&&...) {} };
struct scoped_estream_redirect { template <class... A> scoped_estream_redirect(A&&...) {} };
template <class... T> struct call_guard {};
struct gil_scoped_release {};
struct object {};
struct list : object {};
struct dict : object {};
template <class T> struct array_t {};
} // namespace pybind11
#define PYBIND11_MAKE_OPAQUE(...)

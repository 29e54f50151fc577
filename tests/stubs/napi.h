// napi.h -- TEST STUB, not node-addon-api.  Declarations only, of the part of the documented node-addon-api C++
// surface that fusion_sim_b200/js/fusionsim_napi.cc uses, so that the addon source can at least be type-checked
// (g++ -fsyntax-only) in an image without Node.js.  Nothing here can run; the real header comes from
// `npm install node-addon-api` where Node exists.
#pragma once
#include <cstddef>
#include <cstdint>
#include <exception>
#include <initializer_list>
#include <string>

namespace Napi {

class Env;
class Boolean;

class Value {
  public:
    bool IsNumber() const;
    bool IsTypedArray() const;
    bool IsUndefined() const;
    template <typename T> T As() const;
    Boolean ToBoolean() const;
    class Env Env() const;
};

class Env {
  public:
    Value Undefined() const;
    Value Null() const;
};

class Boolean : public Value {
  public:
    bool Value() const;
    operator bool() const;
};

class Number : public Value {
  public:
    operator float() const;
    operator double() const;
    operator int32_t() const;
    operator uint32_t() const;
    operator int64_t() const;
    double DoubleValue() const;
    float FloatValue() const;
    int32_t Int32Value() const;
    uint32_t Uint32Value() const;
    int64_t Int64Value() const;
};

class String : public Value {
  public:
    operator std::string() const;
    std::string Utf8Value() const;
};

class Object : public Value {
  public:
    Value Get(const char *key) const;
    Value Get(const std::string &key) const;
    bool Has(const char *key) const;
    bool Has(const std::string &key) const;
    void Set(const char *key, const Value &value);
    void Set(const std::string &key, const Value &value);
};

class Function : public Object {};

template <typename T>
class TypedArrayOf : public Object {
  public:
    T *Data();
    const T *Data() const;
    size_t ElementLength() const;
};
using Float64Array = TypedArrayOf<double>;
using Float32Array = TypedArrayOf<float>;
using Uint8Array = TypedArrayOf<uint8_t>;

class CallbackInfo {
  public:
    class Env Env() const;
    size_t Length() const;
    const Value operator[](size_t index) const;
    Value This() const;
};

class Error : public std::exception {
  public:
    static Error New(class Env env, const char *message);
    static Error New(class Env env, const std::string &message);
    const char *what() const noexcept override;
};

class PropertyDescriptorStub {};

template <typename T>
class ObjectWrap {
  public:
    explicit ObjectWrap(const CallbackInfo &info);
    virtual ~ObjectWrap();
    using InstanceMethodCallback = Value (T::*)(const CallbackInfo &info);
    using PropertyDescriptor = PropertyDescriptorStub;
    static PropertyDescriptor InstanceMethod(const char *utf8name, InstanceMethodCallback method);
    static Function DefineClass(class Env env, const char *utf8name, const std::initializer_list<PropertyDescriptor> &properties);
};

}  // namespace Napi

#define NODE_API_MODULE(modname, regfunc) \
    Napi::Object napi_stub_register_##modname(Napi::Env env, Napi::Object exports) { return regfunc(env, exports); }

// Teuchos_shim.hpp — the handful of Teuchos facilities /root/reference/ExodusIO.hpp touches (RCP, Array,
// Comm, OrdinalTraits, ParameterList, VerboseObjectBase), single rank.  TEST INFRASTRUCTURE.
#pragma once
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "mpi.h"

namespace Teuchos {

template <class T>
class RCP {
   public:
    RCP() = default;
    RCP(std::nullptr_t) {}
    explicit RCP(T *p) : p_(p) {}
    RCP(std::shared_ptr<T> p) : p_(std::move(p)) {}
    template <class U, class = typename std::enable_if<std::is_convertible<U *, T *>::value>::type>
    RCP(const RCP<U> &o) : p_(o.shared()) {}
    T *operator->() const { return p_.get(); }
    T &operator*() const { return *p_; }
    T *get() const { return p_.get(); }
    bool is_null() const { return !p_; }
    RCP<const T> getConst() const { return RCP<const T>(std::shared_ptr<const T>(p_)); }
    const std::shared_ptr<T> &shared() const { return p_; }

   private:
    std::shared_ptr<T> p_;
};
template <class T>
RCP<T> rcp(T *p) { return RCP<T>(p); }

template <class T>
class Array : public std::vector<T> {
   public:
    using std::vector<T>::vector;
    int size() const { return (int)std::vector<T>::size(); }      // Teuchos::Array::size() is signed
};
template <class T>
using ArrayView = Array<T>;
template <class T>
using ArrayRCP = Array<T>;

template <class T>
struct OrdinalTraits {
    static T invalid() { return std::numeric_limits<T>::is_signed ? (T)-1 : std::numeric_limits<T>::max(); }
};

template <class T>
struct ScalarTraits {
    typedef T magnitudeType;
    static T zero() { return T(0); }
    static T one() { return T(1); }
};

// only constructed by the reference's main()s, which the driver does not call
class CommandLineProcessor {
   public:
    CommandLineProcessor(bool = true, bool = true) {}
    template <class T> void setOption(const char *, T *, const char * = "", bool = false) {}
    void setOption(const char *, const char *, bool *, const char * = "") {}
    int parse(int, char **) { return 0; }
};

template <class Ordinal = int>
class Comm {
   public:
    int getRank() const { return 0; }
    int getSize() const { return 1; }
    void barrier() const {}
};
template <class O> int rank(const Comm<O> &) { return 0; }
template <class O> int size(const Comm<O> &) { return 1; }
template <class O> void barrier(const Comm<O> &) {}

enum EVerbosityLevel { VERB_DEFAULT = -1, VERB_NONE = 0, VERB_LOW, VERB_MEDIUM, VERB_HIGH, VERB_EXTREME };

struct FancyOStream : std::ostream {
    FancyOStream() : std::ostream(std::cout.rdbuf()) {}
};
struct VerboseObjectBase {
    static RCP<FancyOStream> getDefaultOStream() { return rcp(new FancyOStream()); }
};

class ParameterList {
   public:
    template <class T> ParameterList &set(const std::string &name, const T &) { names_.push_back(name); return *this; }
    ParameterList &set(const std::string &name, const char *) { names_.push_back(name); return *this; }
   private:
    std::vector<std::string> names_;
};

}  // namespace Teuchos

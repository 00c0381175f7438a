// ompl_min: ompl::base::GenericParam / SpecificParam / ParamSet (string-typed planner params).
#ifndef OMPL_MIN_BASE_GENERICPARAM_H
#define OMPL_MIN_BASE_GENERICPARAM_H
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "ompl/util/Console.h"
namespace ompl {
namespace base {
class GenericParam {
public:
  explicit GenericParam(const std::string &name) : name_(name) {}
  virtual ~GenericParam() {}
  const std::string &getName() const { return name_; }
  virtual bool setValue(const std::string &value) = 0;
  virtual std::string getValue() const = 0;
  void setRangeSuggestion(const std::string &r) { rangeSuggestion_ = r; }
  const std::string &getRangeSuggestion() const { return rangeSuggestion_; }

protected:
  std::string name_;
  std::string rangeSuggestion_;
};
typedef std::shared_ptr<GenericParam> GenericParamPtr;

template <typename T> class SpecificParam : public GenericParam {
public:
  typedef std::function<void(T)> SetterFn;
  typedef std::function<T()> GetterFn;
  SpecificParam(const std::string &name, const SetterFn &setter, const GetterFn &getter = GetterFn())
      : GenericParam(name), setter_(setter), getter_(getter) {}
  bool setValue(const std::string &value) override {
    bool ok = true;
    try {
      if (setter_)
        setter_(parse(value, (T *)nullptr));
    } catch (std::exception &e) {
      ok = false;
      OMPL_WARN("Invalid value format specified for parameter '%s': %s", name_.c_str(), e.what());
    }
    return ok;
  }
  std::string getValue() const override {
    if (!getter_)
      return "";
    std::ostringstream ss;
    ss << getter_();
    return ss.str();
  }

private:
  template <typename U> static U parse(const std::string &v, U *) {
    std::istringstream ss(v);
    U out;
    ss >> out;
    if (ss.fail())
      throw std::runtime_error("cannot parse '" + v + "'");
    return out;
  }
  static std::string parse(const std::string &v, std::string *) { return v; }
  SetterFn setter_;
  GetterFn getter_;
};

class ParamSet {
public:
  template <typename T>
  void declareParam(const std::string &name, const typename SpecificParam<T>::SetterFn &setter,
                    const typename SpecificParam<T>::GetterFn &getter =
                        typename SpecificParam<T>::GetterFn()) {
    params_[name].reset(new SpecificParam<T>(name, setter, getter));
  }
  bool setParam(const std::string &key, const std::string &value) {
    auto it = params_.find(key);
    if (it == params_.end()) {
      OMPL_ERROR("Parameter '%s' was not found", key.c_str());
      return false;
    }
    return it->second->setValue(value);
  }
  bool getParam(const std::string &key, std::string &value) const {
    auto it = params_.find(key);
    if (it == params_.end())
      return false;
    value = it->second->getValue();
    return true;
  }
  bool hasParam(const std::string &key) const { return params_.count(key) != 0; }
  const GenericParamPtr &getParam(const std::string &key) const {
    static GenericParamPtr empty;
    auto it = params_.find(key);
    return it == params_.end() ? empty : it->second;
  }
  const GenericParamPtr &operator[](const std::string &key) const { return getParam(key); }
  void getParamNames(std::vector<std::string> &names) const {
    names.clear();
    for (auto &p : params_)
      names.push_back(p.first);
  }
  std::size_t size() const { return params_.size(); }

private:
  std::map<std::string, GenericParamPtr> params_;
};
} // namespace base
} // namespace ompl
#endif

/**
 * \file env.h
 * \brief Process-wide configuration lookup: an in-process override map consulted
 *        before getenv(). All DMLC_* / PS_* / BYTEPS_* knobs go through here so tests
 *        can configure several logical nodes inside one process.
 * Parity: reference include/ps/internal/env.h:15-63.
 */
#ifndef PS_INTERNAL_ENV_H_
#define PS_INTERNAL_ENV_H_
#include <cstdlib>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace ps {

class Environment {
 public:
  /*! \brief the singleton */
  static Environment* Get() { return Shared(nullptr).get(); }
  /*! \brief keep the singleton alive past static destruction order */
  static std::shared_ptr<Environment> _GetSharedRef() { return Shared(nullptr); }
  /*! \brief install overrides (first call wins, like the reference's Init) */
  static Environment* Init(const std::unordered_map<std::string, std::string>& envs) {
    return Shared(&envs).get();
  }
  /*! \brief set or replace one override; thread-safe */
  void set(const std::string& k, const std::string& v) {
    std::lock_guard<std::mutex> lk(mu_);
    // values are never erased so pointers handed out by find() stay valid
    auto it = kvs_.find(k);
    if (it == kvs_.end()) {
      kvs_.emplace(k, std::unique_ptr<std::string>(new std::string(v)));
    } else {
      graveyard_.push_back(std::move(it->second));
      it->second.reset(new std::string(v));
    }
  }
  /*! \brief override map first, then the process environment; nullptr if unset */
  const char* find(const char* k) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = kvs_.find(k);
      if (it != kvs_.end()) return it->second->c_str();
    }
    return getenv(k);
  }

 private:
  Environment() {}
  static std::shared_ptr<Environment> Shared(
      const std::unordered_map<std::string, std::string>* envs) {
    static std::shared_ptr<Environment> inst(new Environment());
    if (envs) {
      for (const auto& kv : *envs) inst->set(kv.first, kv.second);
    }
    return inst;
  }
  std::mutex mu_;
  std::unordered_map<std::string, std::unique_ptr<std::string>> kvs_;
  std::vector<std::unique_ptr<std::string>> graveyard_;
};

}  // namespace ps
#endif  // PS_INTERNAL_ENV_H_

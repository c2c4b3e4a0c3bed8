// sb200/matcher.h -- dispatch of a preprocessing call to the implementation function that
// matches the input formats, converting inputs when allowed.
//
// Interface mirrored: utils::FunctionMatcherMixin, utils/function_matcher_mixin.h:35-416
//   RegisterFunction / RegisterFunctionNoOverride / UnregisterFunction   (:59-75, :282-300)
//   Execute (deletes converted intermediates)                              (:228-245)
//   CachedExecute (returns them)                                           (:171-226)
//   GetFunction: exact key + every input in an allowed context -> direct call; otherwise the
//   registered key reachable with the cheapest conversion schema (unit cost per hop); an input
//   whose type already equals the key's costs nothing even if its context is not in the list
//   (the reference's rule, :366-369); nothing usable -> FunctionNotFoundException (:389-405);
//   conversion needed but convert_input == false -> DirectExecutionNotAvailableException
//   (:196-202).
#pragma once
#include <limits>
#include <unordered_map>

#include "format.h"

namespace sparsebase::utils {

template <typename ReturnType>
using PreprocessFunction = ReturnType (*)(std::vector<format::Format *> formats,
                                          utils::Parameters *params);

template <typename ReturnType, class PreprocessingImpl = Parameterizable,
          typename Function = PreprocessFunction<ReturnType>,
          typename Key = std::vector<std::type_index>, typename KeyHash = TypeIndexVectorHash,
          typename KeyEqualTo = std::equal_to<std::vector<std::type_index>>>
class FunctionMatcherMixin : public PreprocessingImpl {
  typedef std::unordered_map<Key, Function, KeyHash, KeyEqualTo> FunctionMap;

 public:
  std::vector<Key> GetAvailableFormats() {
    std::vector<Key> keys;
    for (const auto &kv : map_to_function_) keys.push_back(kv.first);
    return keys;
  }
  bool RegisterFunctionNoOverride(const Key &key, const Function &fn) {
    return map_to_function_.emplace(key, fn).second;
  }
  void RegisterFunction(const Key &key, const Function &fn) { map_to_function_[key] = fn; }
  bool UnregisterFunction(const Key &key) { return map_to_function_.erase(key) > 0; }

 protected:
  using PreprocessingImpl::PreprocessingImpl;
  FunctionMap map_to_function_;

  bool CheckIfKeyMatches(const FunctionMap &map, const Key &key,
                         const std::vector<format::Format *> &inputs,
                         const std::vector<context::Context *> &contexts) {
    if (map.find(key) == map.end()) return false;
    for (format::Format *f : inputs) {
      bool ok = false;
      for (context::Context *c : contexts) ok = ok || f->get_context()->IsEquivalent(c);
      if (!ok) return false;
    }
    return true;
  }

  std::tuple<Function, converter::ConversionSchema> GetFunction(
      const std::vector<format::Format *> &inputs, const Key &key, const FunctionMap &map,
      const std::vector<context::Context *> &contexts) {
    if (CheckIfKeyMatches(map, key, inputs, contexts))
      return std::make_tuple(map.at(key), converter::ConversionSchema(key.size()));
    bool found = false;
    unsigned best_cost = std::numeric_limits<unsigned>::max();
    Function best_fn = nullptr;
    converter::ConversionSchema best_schema;
    for (const auto &candidate : map) {
      const Key &want = candidate.first;
      if (want.size() != key.size()) continue;
      converter::ConversionSchema schema;
      unsigned cost = 0;
      bool usable = true;
      for (size_t i = 0; i < want.size() && usable; i++) {
        if (key[i] == want[i]) {
          schema.push_back({});
          continue;
        }
        auto conv = inputs[i]->get_converter();
        if (!conv) throw utils::NoConverterException();
        converter::ConversionChain chain =
            conv->GetConversionChain(key[i], inputs[i]->get_context(), want[i], contexts);
        if (!chain) {
          usable = false;
          break;
        }
        cost += std::get<1>(*chain);
        schema.push_back(std::move(chain));
      }
      if (usable && (!found || cost < best_cost)) {
        found = true;
        best_cost = cost;
        best_fn = candidate.second;
        best_schema = std::move(schema);
      }
    }
    if (!found) {
      std::string msg = "Could not find a function that matches the formats: {";
      for (format::Format *f : inputs) msg += f->get_name() + " ";
      msg += "} using the contexts {";
      for (context::Context *c : contexts) msg += c->get_name() + " ";
      throw utils::FunctionNotFoundException(msg + "}");
    }
    return std::make_tuple(best_fn, best_schema);
  }

  template <typename F, typename... SF>
  std::tuple<std::vector<std::vector<format::Format *>>, ReturnType> CachedExecute(
      utils::Parameters *params, std::vector<context::Context *> contexts, bool convert_input,
      bool clear_intermediate, F first, SF... rest) {
    std::vector<format::Format *> inputs{first, rest...};
    Key key;
    for (format::Format *f : inputs) key.push_back(f->get_id());
    auto [fn, schema] = GetFunction(inputs, key, map_to_function_, contexts);
    if (!convert_input)
      for (const auto &chain : schema)
        if (chain)
          throw utils::DirectExecutionNotAvailableException<Key>(key, GetAvailableFormats());
    std::vector<std::vector<format::Format *>> chains =
        converter::Converter::ApplyConversionSchema(schema, inputs, clear_intermediate);
    std::vector<format::Format *> finals;
    std::vector<std::vector<format::Format *>> created;
    for (const auto &c : chains) {
      finals.push_back(c.back());
      created.emplace_back(c.begin() + 1, c.end());
    }
    return std::make_tuple(created, fn(finals, params));
  }

  template <typename F, typename... SF>
  ReturnType Execute(utils::Parameters *params, std::vector<context::Context *> contexts,
                     bool convert_input, F first, SF... rest) {
    auto out = CachedExecute(params, std::move(contexts), convert_input, true, first, rest...);
    for (auto &chain : std::get<0>(out))
      for (format::Format *f : chain) delete f;
    return std::get<1>(out);
  }
};

}  // namespace sparsebase::utils

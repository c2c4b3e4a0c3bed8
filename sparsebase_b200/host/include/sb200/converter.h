// sb200/converter.h -- format::Format (the polymorphic base of every container) and the
// conversion-function registry with its chain finder.
//
// Interfaces mirrored (reference paths relative to src/sparsebase/):
//   format::Format, Ownership, DimensionType      format/format.h:41-163
//   converter::ConversionFunction / Condition     converter/converter.h:42-48
//   converter::Converter                          converter/converter.h:72-350,
//                                                 converter/converter.cc:35-287
//   converter::ConverterStore                     converter/converter_store.h:10-41
//
// Behaviour kept: two maps (copy / move conversions); several (condition, function) pairs per
// (from, to) edge, the first registered pair whose condition holds for one of the allowed
// contexts wins; breadth-first search with unit edge cost; ConvertCached returns the source
// untouched when it already has the requested type in an equivalent context; a missing chain
// raises utils::ConversionException.
// Deliberate fix (SURVEY.md App. A spirit): while searching a multi-hop chain the reference
// evaluates every edge condition against the SOURCE's context (converter.cc:166); here the
// context reached by the previous hop is used, so host -> device -> device -> host chains are
// found.  Single-hop behaviour is identical.
#pragma once
#include <algorithm>
#include <deque>
#include <functional>
#include <mutex>
#include <optional>
#include <tuple>
#include <unordered_map>

#include "core.h"

namespace sparsebase {

namespace converter {
class Converter;
}

// ======================================================================= format::Format
namespace format {

enum Ownership { kNotOwned = 0, kOwned = 1 };

typedef unsigned long long DimensionType;

class Format : public utils::Identifiable {
 public:
  virtual ~Format() = default;
  virtual Format *Clone() const = 0;
  virtual std::vector<DimensionType> get_dimensions() const = 0;
  virtual DimensionType get_num_nnz() const = 0;
  virtual DimensionType get_order() const = 0;
  virtual context::Context *get_context() const = 0;
  virtual std::shared_ptr<converter::Converter const> get_converter() const = 0;

  // Cast to the concrete class; utils::TypeException when the object is something else.
  template <typename T>
  typename std::remove_pointer<T>::type *AsAbsolute() {
    using TBase = typename std::remove_pointer<T>::type;
    static_assert(std::is_base_of_v<Format, TBase>,
                  "Cannot cast a non-Format class using AsAbsolute");
    if (this->get_id() == std::type_index(typeid(TBase))) return static_cast<TBase *>(this);
    throw utils::TypeException(get_name(), utils::demangle(typeid(TBase).name()));
  }
  template <typename T>
  bool IsAbsolute() {
    using TBase = typename std::remove_pointer<T>::type;
    return this->get_id() == std::type_index(typeid(TBase));
  }
};

}  // namespace format

// ======================================================================= converter
namespace converter {

typedef std::function<format::Format *(format::Format *, context::Context *)> ConversionFunction;
typedef std::function<bool(context::Context *, context::Context *)> ConversionCondition;
typedef std::tuple<ConversionFunction, context::Context *, utils::CostType> ConversionStep;
typedef std::optional<std::tuple<std::vector<ConversionStep>, utils::CostType>> ConversionChain;
typedef std::vector<ConversionChain> ConversionSchema;
typedef std::unordered_map<
    std::type_index,
    std::unordered_map<std::type_index,
                       std::vector<std::tuple<ConversionCondition, ConversionFunction>>>>
    ConversionMap;

class Converter {
 public:
  virtual ~Converter() = default;
  virtual std::type_index get_converter_type() const = 0;
  virtual Converter *Clone() const = 0;
  virtual void Reset() = 0;

  void RegisterConversionFunction(std::type_index from_type, std::type_index to_type,
                                  ConversionFunction conv_func,
                                  ConversionCondition edge_condition,
                                  bool is_move_conversion = false) {
    map_for(is_move_conversion)[from_type][to_type].emplace_back(std::move(edge_condition),
                                                                 std::move(conv_func));
  }

  void ClearConversionFunctions(std::type_index from_type, std::type_index to_type,
                                bool move_conversion = false) {
    auto &map = map_for(move_conversion);
    auto it = map.find(from_type);
    if (it == map.end()) return;
    it->second.erase(to_type);
    if (it->second.empty()) map.erase(it);
  }
  void ClearConversionFunctions(bool move_conversion = false) { map_for(move_conversion).clear(); }

  ConversionChain GetConversionChain(std::type_index from_type, context::Context *from_context,
                                     std::type_index to_type,
                                     const std::vector<context::Context *> &to_contexts,
                                     bool is_move_conversion = false) const {
    if (from_type == to_type &&
        std::find(to_contexts.begin(), to_contexts.end(), from_context) != to_contexts.end())
      return ConversionChain(std::in_place);  // nothing to do: an empty but present chain
    std::vector<ConversionStep> steps =
        Search(from_type, from_context, to_type, to_contexts, map_for(is_move_conversion));
    if (steps.empty()) return {};
    const utils::CostType cost = (utils::CostType)steps.size();
    return std::make_tuple(std::move(steps), cost);
  }

  bool CanConvert(std::type_index from_type, context::Context *from_context,
                  std::type_index to_type, context::Context *to_context,
                  bool is_move_conversion = false) const {
    return GetConversionChain(from_type, from_context, to_type, {to_context}, is_move_conversion)
        .has_value();
  }
  bool CanConvert(std::type_index from_type, context::Context *from_context,
                  std::type_index to_type, const std::vector<context::Context *> &to_contexts,
                  bool is_move_conversion = false) const {
    return GetConversionChain(from_type, from_context, to_type, to_contexts, is_move_conversion)
        .has_value();
  }

  // Every format produced along the chain (the source itself when no conversion is needed).
  std::vector<format::Format *> ConvertCached(format::Format *source, std::type_index to_type,
                                              std::vector<context::Context *> to_contexts,
                                              bool is_move_conversion = false) const {
    if (to_type == source->get_id())
      for (context::Context *c : to_contexts)
        if (c->IsEquivalent(source->get_context())) return {source};
    ConversionChain chain = GetConversionChain(source->get_id(), source->get_context(), to_type,
                                               to_contexts, is_move_conversion);
    if (!chain) throw utils::ConversionException(source->get_name(), utils::demangle(to_type));
    std::vector<format::Format *> all = ApplyConversionChain(chain, source, false);
    return std::vector<format::Format *>(all.begin() + 1, all.end());
  }
  std::vector<format::Format *> ConvertCached(format::Format *source, std::type_index to_type,
                                              context::Context *to_context,
                                              bool is_move_conversion = false) const {
    return ConvertCached(source, to_type, std::vector<context::Context *>{to_context},
                         is_move_conversion);
  }

  // The final format only; intermediates of a multi-hop chain are deleted.
  format::Format *Convert(format::Format *source, std::type_index to_type,
                          std::vector<context::Context *> to_contexts,
                          bool is_move_conversion = false) const {
    std::vector<format::Format *> outs =
        ConvertCached(source, to_type, std::move(to_contexts), is_move_conversion);
    for (size_t i = 0; i + 1 < outs.size(); i++) delete outs[i];
    return outs.back();
  }
  format::Format *Convert(format::Format *source, std::type_index to_type,
                          context::Context *to_context, bool is_move_conversion = false) const {
    return Convert(source, to_type, std::vector<context::Context *>{to_context},
                   is_move_conversion);
  }
  template <typename FormatType>
  FormatType *Convert(format::Format *source, context::Context *to_context,
                      bool is_move_conversion = false) const {
    return Convert(source, FormatType::get_id_static(), to_context, is_move_conversion)
        ->template AsAbsolute<FormatType>();
  }
  template <typename FormatType>
  FormatType *Convert(format::Format *source, std::vector<context::Context *> to_contexts,
                      bool is_move_conversion = false) const {
    return Convert(source, FormatType::get_id_static(), std::move(to_contexts),
                   is_move_conversion)
        ->template AsAbsolute<FormatType>();
  }

  // chain[0] = input, then one entry per step (only the last one when clear_intermediate).
  static std::vector<format::Format *> ApplyConversionChain(const ConversionChain &chain,
                                                            format::Format *input,
                                                            bool clear_intermediate) {
    std::vector<format::Format *> produced{input};
    if (!chain) return produced;
    const std::vector<ConversionStep> &steps = std::get<0>(*chain);
    format::Format *cur = input;
    for (size_t i = 0; i < steps.size(); i++) {
      format::Format *next = std::get<0>(steps[i])(cur, std::get<1>(steps[i]));
      const bool last = i + 1 == steps.size();
      if (!clear_intermediate || last) produced.push_back(next);
      if (clear_intermediate && i != 0) delete cur;  // an intermediate this call created
      cur = next;
    }
    return produced;
  }

  static std::vector<std::vector<format::Format *>> ApplyConversionSchema(
      const ConversionSchema &cs, const std::vector<format::Format *> &packed_sfs,
      bool clear_intermediate) {
    std::vector<std::vector<format::Format *>> out;
    for (size_t i = 0; i < cs.size(); i++)
      out.push_back(ApplyConversionChain(cs[i], packed_sfs[i], clear_intermediate));
    return out;
  }

 protected:
  ConversionMap copy_conversion_map_;
  ConversionMap move_conversion_map_;

 private:
  ConversionMap &map_for(bool move) { return move ? move_conversion_map_ : copy_conversion_map_; }
  const ConversionMap &map_for(bool move) const {
    return move ? move_conversion_map_ : copy_conversion_map_;
  }

  // Breadth-first search over format types; every node remembers the context it is reached in.
  static std::vector<ConversionStep> Search(std::type_index from_type,
                                            context::Context *from_context,
                                            std::type_index to_type,
                                            const std::vector<context::Context *> &to_contexts,
                                            const ConversionMap &map) {
    struct Reached {
      std::type_index parent;
      ConversionStep step;
      context::Context *ctx;
    };
    std::unordered_map<std::type_index, Reached> reached;
    reached.emplace(from_type,
                    Reached{from_type, ConversionStep{nullptr, nullptr, 0}, from_context});
    std::deque<std::type_index> frontier{from_type};
    while (!frontier.empty()) {
      const std::type_index cur = frontier.front();
      frontier.pop_front();
      auto edges = map.find(cur);
      if (edges == map.end()) continue;
      context::Context *cur_ctx = reached.at(cur).ctx;
      for (const auto &nb : edges->second) {
        if (reached.count(nb.first)) continue;
        bool taken = false;
        for (const auto &cond_fn : nb.second) {  // registration order: first usable pair wins
          for (context::Context *to_ctx : to_contexts) {
            if (!std::get<0>(cond_fn)(cur_ctx, to_ctx)) continue;
            reached.emplace(nb.first,
                            Reached{cur, ConversionStep{std::get<1>(cond_fn), to_ctx, 1}, to_ctx});
            frontier.push_back(nb.first);
            taken = true;
            break;
          }
          if (taken) break;
        }
        if (taken && nb.first == to_type) {
          std::vector<ConversionStep> steps;
          for (std::type_index t = to_type; t != from_type; t = reached.at(t).parent)
            steps.push_back(reached.at(t).step);
          std::reverse(steps.begin(), steps.end());
          return steps;
        }
      }
    }
    return {};
  }
};

template <class ConverterType>
class ConverterImpl : public Converter {
 public:
  std::type_index get_converter_type() const override { return typeid(ConverterType); }
};

// One shared converter instance per converter type, handed to every format of that type
// triple (weak_ptr cache guarded by a mutex, like converter_store.h:20-40).
class ConverterStore {
 public:
  static ConverterStore &GetStore() {
    static ConverterStore store;
    return store;
  }
  template <typename ConverterType>
  std::shared_ptr<ConverterType> get_converter() {
    std::lock_guard<std::mutex> lock(mu_);
    auto it = cache_.find(std::type_index(typeid(ConverterType)));
    if (it != cache_.end())
      if (std::shared_ptr<Converter> live = it->second.lock())
        return std::static_pointer_cast<ConverterType>(live);
    auto fresh = std::make_shared<ConverterType>();
    cache_[std::type_index(typeid(ConverterType))] = fresh;
    return fresh;
  }

 private:
  ConverterStore() = default;
  std::mutex mu_;
  std::unordered_map<std::type_index, std::weak_ptr<Converter>> cache_;
};

}  // namespace converter
}  // namespace sparsebase

// gr4b200 host layer -- tag-aware source / sink fixtures
// (blocks/testing/include/gnuradio-4.0/testing/TagMonitors.hpp: TagSource, TagSink; the subset the filter tests use,
// blocks/filter/test/qa_filter.cpp:267-320).
#pragma once

#include <vector>

#include "../Block.hpp"

namespace gr::testing {

// produces 0, 1, 2, ... (n_samples_max of them) and publishes `_tags` at their absolute sample indices; its
// `sample_rate` (like every changed stream setting) goes out as a tag with the first sample
template<typename T>
struct TagSource : gr::Block<TagSource<T>> {
    using gr::Block<TagSource<T>>::Block;
    gr::PortOut<T>       out;
    gr::Size_t           n_samples_max = 1024;
    float                sample_rate   = 1000.f;
    std::string          signal_name   = "unknown signal";
    GR_MAKE_REFLECTABLE(TagSource, out, n_samples_max, sample_rate, signal_name);
    std::vector<gr::Tag> _tags; // ascending index
    gr::Size_t           _nSamplesProduced = 0;
    std::size_t          _nextTag          = 0;

    gr::work::Status processBulk(std::span<T> output) {
        const std::size_t n = std::min<std::size_t>(output.size(), n_samples_max - _nSamplesProduced);
        for (std::size_t i = 0; i < n; ++i) {
            output[i] = static_cast<T>(static_cast<float>(_nSamplesProduced + i));
        }
        while (_nextTag < _tags.size() && _tags[_nextTag].index < _nSamplesProduced + n) {
            this->publishTag(_tags[_nextTag].map, _tags[_nextTag].index - _nSamplesProduced);
            ++_nextTag;
        }
        _nSamplesProduced += static_cast<gr::Size_t>(n);
        this->publishOnly(n);
        return _nSamplesProduced >= n_samples_max ? gr::work::Status::DONE : gr::work::Status::OK;
    }
};

// records every tag it receives (with the absolute index of the chunk it arrived with); `sample_rate` follows the tags
template<typename T>
struct TagSink : gr::Block<TagSink<T>> {
    using gr::Block<TagSink<T>>::Block;
    gr::PortIn<T>        in;
    gr::Size_t           n_samples_expected = 0;
    float                sample_rate        = 1000.f;
    std::string          signal_name        = "unknown signal";
    GR_MAKE_REFLECTABLE(TagSink, in, n_samples_expected, sample_rate, signal_name);
    std::vector<gr::Tag> _tags;
    std::vector<T>       _samples;
    gr::Size_t           _nSamplesProduced = 0;

    gr::work::Status processBulk(std::span<const T> input) {
        if (this->inputTagsPresent()) {
            _tags.push_back(gr::Tag{_nSamplesProduced, this->mergedInputTag().map});
        }
        _samples.insert(_samples.end(), input.begin(), input.end());
        _nSamplesProduced += static_cast<gr::Size_t>(input.size());
        return n_samples_expected > 0 && _nSamplesProduced >= n_samples_expected ? gr::work::Status::DONE : gr::work::Status::OK;
    }
};

} // namespace gr::testing

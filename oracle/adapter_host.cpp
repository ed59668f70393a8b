// TEST INFRASTRUCTURE ONLY.
//
// Host program that EXECUTES the reference-side binding include/B200ModelRunner.h: it is compiled against the reference's
// own headers (ModelRunnerBase, BasecallModelConfig, crf_utils) and libtorch, exactly as the header would be inside
// dorado, and drives the engine through the reference's interface only:
//   config::load_model_config -> B200Caller (torch::load of the *.tensor files through the reference's
//   load_crf_model_weights) -> B200ModelRunner : ModelRunnerBase -> accept_chunk(at::Tensor) / call_chunks() / sample_stats()
// Built here (where /root/reference exists) by oracle/Makefile into oracle/_ref/adapter_host; run on the GPU box by
// tests/test_zz_adapter_gpu.py, which compares its output with the ctypes path on the same chunks.
//
// usage: adapter_host <config dir> <weights.b2w> <work dir> <batch> <chunk size> <signal.f16> <num chunks> <out file>
#include "B200ModelRunner.h"

#include "basecall/ModelRunnerBase.h"
#include "config/BasecallModelConfig.h"

#include <torch/serialize.h>
#include <torch/torch.h>

#include <cstdio>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

namespace {

void write_tensor_files(const std::string& b2w, const std::filesystem::path& dst) {
    std::ifstream f(b2w, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + b2w);
    char magic[4];
    uint32_t n = 0;
    f.read(magic, 4);
    f.read(reinterpret_cast<char*>(&n), 4);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t name_len = 0, ndim = 0;
        f.read(reinterpret_cast<char*>(&name_len), 4);
        std::string name(name_len, '\0');
        f.read(name.data(), name_len);
        f.read(reinterpret_cast<char*>(&ndim), 4);
        std::vector<int64_t> dims(ndim);
        int64_t numel = 1;
        for (uint32_t d = 0; d < ndim; ++d) {
            uint32_t v = 0;
            f.read(reinterpret_cast<char*>(&v), 4);
            dims[d] = v;
            numel *= v;
        }
        at::Tensor t = at::empty(dims, at::kFloat);
        f.read(reinterpret_cast<char*>(t.data_ptr<float>()), numel * 4);
        if (!f) throw std::runtime_error("truncated weights file");
        torch::save(std::vector<at::Tensor>{t}, (dst / name).string());
    }
}

}  // namespace

int main(int argc, char** argv) {
    if (argc != 9) {
        std::fprintf(stderr, "usage: %s <config dir> <weights.b2w> <work dir> <batch> <chunk size> <signal.f16> <num chunks> <out>\n", argv[0]);
        return 2;
    }
    try {
        namespace fs = std::filesystem;
        at::InferenceMode guard;
        const fs::path src(argv[1]);
        const fs::path dst = fs::path(argv[3]) / src.filename();
        fs::create_directories(dst);
        fs::copy_file(src / "config.toml", dst / "config.toml", fs::copy_options::overwrite_existing);
        write_tensor_files(argv[2], dst);
        const int batch = std::stoi(argv[4]), chunk = std::stoi(argv[5]), n = std::stoi(argv[7]);

        auto config = dorado::config::load_model_config(dst);
        config.basecaller.set_batch_size(batch);
        config.basecaller.set_chunk_size(chunk);
        config.normalise_basecaller_params();

        auto caller = std::make_shared<dorado::basecall::B200Caller>(config, 0);
        // held through the reference's own base-class pointer: only ModelRunnerBase's interface is used from here on
        dorado::basecall::RunnerPtr runner =
                std::make_unique<dorado::basecall::B200ModelRunner>(caller, batch, config.basecaller.chunk_size());
        const int64_t T = static_cast<int64_t>(runner->chunk_size());

        std::vector<uint16_t> sig(static_cast<size_t>(n) * T);
        {
            std::ifstream f(argv[6], std::ios::binary);
            f.read(reinterpret_cast<char*>(sig.data()), sig.size() * 2);
            if (!f) throw std::runtime_error("signal file too short");
        }
        std::ofstream out(argv[8]);
        out << "name " << runner->get_name() << "\n";
        out << "dims " << runner->batch_size() << " " << runner->chunk_size() << " " << runner->config().stride << "\n";
        const auto to = runner->batch_timeouts_ms();
        out << "timeouts " << to.first << " " << to.second << " low_latency " << runner->is_low_latency() << " vcs "
            << runner->variable_chunk_sizes() << "\n";
        for (int start = 0; start < n; start += batch) {
            const int cnt = std::min(batch, n - start);
            for (int i = 0; i < cnt; ++i) {
                // [C_in = 1, chunk_size] half tensor, as BasecallerNode hands it over (BasecallerNode.cpp:443-444)
                at::Tensor c = at::from_blob(sig.data() + static_cast<size_t>(start + i) * T, {1, T}, at::kHalf);
                runner->accept_chunk(i, c);
            }
            const std::vector<dorado::basecall::decode::DecodedChunk> res = runner->call_chunks(cnt);
            if (static_cast<int>(res.size()) != cnt) throw std::runtime_error("call_chunks returned the wrong count");
            for (const auto& r : res) {
                std::string mv(r.moves.size(), '0');
                for (size_t k = 0; k < r.moves.size(); ++k) mv[k] = r.moves[k] ? '1' : '0';
                out << "chunk " << r.sequence << " " << r.qstring << " " << mv << "\n";
            }
        }
        const auto stats = runner->sample_stats();
        out << "stats batches_called " << stats.at("batches_called") << " model_decode_ms " << stats.at("model_decode_ms") << "\n";
        runner->terminate();
        bool refused = false;
        try {
            runner->call_chunks(1);
        } catch (const std::exception&) {
            refused = true;  // errors cross the boundary as exceptions, as the reference's runners throw
        }
        runner->restart();
        out << "terminate_refuses " << refused << " restart_ok " << (runner->call_chunks(1).size() == 1) << "\n";
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "adapter_host: %s\n", e.what());
        return 1;
    }
}

// Multi-device runner pool: the B200 counterpart of api::create_basecall_runners (dorado/api/runner_creation.cpp:46-130)
// plus the worker loop BasecallerNode runs per runner (dorado/read_pipeline/nodes/BasecallerNode.cpp:300-352).
//
// One process, one Engine per device (the reference: one CudaCaller per device), `runners_per_device` Runners per engine
// (the reference's num_runners, default 2), one host thread per runner.  Chunks shard embarrassingly: the worker threads
// take batches from ONE shared cursor, so a faster or less loaded GPU simply takes more batches (the reference's shared
// chunk queues do the same); there is no collective and no inter-GPU traffic.
//
// NUMA: a worker pins itself to the CPUs of its device's host NUMA node (cudaDevAttrHostNumaId) BEFORE it creates its
// runner, so the runner's pinned staging buffers are first-touched on the socket the GPU hangs off, and the accept_chunk
// memcpys and result copies run there too.
#include "engine.h"

#include <pthread.h>
#include <sched.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <thread>

namespace b200 {

namespace {

// "0-31,64-95" -> cpu set
bool parse_cpulist(const std::string& s, cpu_set_t* set) {
    CPU_ZERO(set);
    std::stringstream ss(s);
    std::string tok;
    bool any = false;
    while (std::getline(ss, tok, ',')) {
        int a = 0, b = 0;
        if (std::sscanf(tok.c_str(), "%d-%d", &a, &b) == 2) {
            for (int c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET(c, set), any = true;
        } else if (std::sscanf(tok.c_str(), "%d", &a) == 1) {
            if (a < CPU_SETSIZE) CPU_SET(a, set), any = true;
        }
    }
    return any;
}

int pin_to_device_numa_node(int device) {
    int node = -1;
    if (cudaDeviceGetAttribute(&node, cudaDevAttrHostNumaId, device) != cudaSuccess || node < 0) {
        cudaGetLastError();
        return -1;
    }
    std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string line;
    cpu_set_t set;
    if (!f || !std::getline(f, line) || !parse_cpulist(line, &set)) return -1;
    if (pthread_setaffinity_np(pthread_self(), sizeof(set), &set) != 0) return -1;
    return node;
}

}  // namespace

struct Pool::Worker {
    int device = 0;
    int numa_node = -1;
    Engine* engine = nullptr;
    std::unique_ptr<Runner> runner;
    std::thread thread;
    std::string error;
    int64_t batches = 0, chunks = 0;
};

Pool::Pool(const b200_model_desc& desc, const b200_tensor* tensors, int num_tensors, const int* devices, int num_devices,
           int runners_per_device, int batch_size, int chunk_size)
        : m_batch(batch_size), m_chunk(chunk_size) {
    if (num_devices < 1 || runners_per_device < 1 || !devices) throw std::invalid_argument("pool: need >= 1 device and runner");
    for (int d = 0; d < num_devices; ++d) {
        m_engines.emplace_back(new Engine(desc, tensors, num_tensors, devices[d]));
        m_engines.back()->set_num_runners(runners_per_device);
    }
    m_t_out = chunk_size / desc.stride;
    for (int d = 0; d < num_devices; ++d) {
        for (int r = 0; r < runners_per_device; ++r) {
            auto w = std::make_unique<Worker>();
            w->device = devices[d];
            w->engine = m_engines[d].get();
            m_workers.push_back(std::move(w));
        }
    }
    // runners are built by their own (pinned) threads; construction errors are collected and rethrown here
    m_ready = 0;
    for (auto& w : m_workers) w->thread = std::thread([this, wp = w.get()] { worker_main(*wp); });
    {
        std::unique_lock<std::mutex> lock(m_mutex);
        m_cv_done.wait(lock, [&] { return m_ready == (int)m_workers.size(); });
    }
    std::string err;
    for (auto& w : m_workers) {
        if (!w->error.empty()) err = w->error;
    }
    if (!err.empty()) {
        shutdown();
        throw std::runtime_error("pool: " + err);
    }
}

Pool::~Pool() { shutdown(); }

void Pool::shutdown() {
    {
        std::lock_guard<std::mutex> lock(m_mutex);
        m_stop = true;
        ++m_job_id;
    }
    m_cv_job.notify_all();
    for (auto& w : m_workers) {
        if (w->thread.joinable()) w->thread.join();
    }
    m_workers.clear();   // runners before their engines
    m_engines.clear();
}

int Pool::num_runners() const { return (int)m_workers.size(); }
Runner* Pool::runner(int i) { return (i >= 0 && i < (int)m_workers.size()) ? m_workers[i]->runner.get() : nullptr; }
int Pool::runner_numa_node(int i) const { return (i >= 0 && i < (int)m_workers.size()) ? m_workers[i]->numa_node : -1; }
int64_t Pool::runner_batches(int i) const { return (i >= 0 && i < (int)m_workers.size()) ? m_workers[i]->batches : 0; }

void Pool::worker_main(Worker& w) {
    try {
        B200_CUDA(cudaSetDevice(w.device));
        w.numa_node = pin_to_device_numa_node(w.device);
        w.runner = std::make_unique<Runner>(*w.engine, m_batch, m_chunk);
    } catch (const std::exception& e) {
        w.error = e.what();
    }
    long long seen = 0;
    {
        std::lock_guard<std::mutex> lock(m_mutex);
        ++m_ready;
    }
    m_cv_done.notify_all();
    for (;;) {
        {
            std::unique_lock<std::mutex> lock(m_mutex);
            m_cv_job.wait(lock, [&] { return m_job_id != seen; });
            seen = m_job_id;
            if (m_stop) return;
        }
        if (w.runner && w.error.empty()) {
            try {
                for (;;) {
                    const int64_t start = m_cursor.fetch_add(m_batch);
                    if (start >= m_job.num_chunks) break;
                    const int cnt = (int)std::min<int64_t>(m_batch, m_job.num_chunks - start);
                    for (int i = 0; i < cnt; ++i) {
                        w.runner->accept_chunk_f16(i, m_job.chunks + (size_t)(start + i) * m_chunk, m_chunk);
                    }
                    const b200_result r = w.runner->call_chunks(cnt);
                    const size_t T = (size_t)r.t_out;
                    if (m_job.moves) std::memcpy(m_job.moves + (size_t)start * T, r.moves, (size_t)cnt * T);
                    if (m_job.n_bases) std::memcpy(m_job.n_bases + start, r.n_bases, (size_t)cnt * sizeof(int32_t));
                    for (int i = 0; i < cnt; ++i) {
                        const size_t nb = (size_t)r.n_bases[i];
                        if (m_job.sequence) std::memcpy(m_job.sequence + (size_t)(start + i) * T, r.sequence + (size_t)i * T, nb);
                        if (m_job.qstring) std::memcpy(m_job.qstring + (size_t)(start + i) * T, r.qstring + (size_t)i * T, nb);
                    }
                    ++w.batches;
                    w.chunks += cnt;
                }
            } catch (const std::exception& e) {
                w.error = e.what();
            }
        }
        {
            std::lock_guard<std::mutex> lock(m_mutex);
            ++m_done;
        }
        m_cv_done.notify_all();
    }
}

double Pool::call_chunks(const uint16_t* chunks, int64_t num_chunks, uint8_t* moves, char* sequence, char* qstring,
                         int32_t* n_bases) {
    if (!chunks || num_chunks < 1) throw std::invalid_argument("pool: no chunks");
    std::lock_guard<std::mutex> one_job(m_job_mutex);
    m_job = Job{chunks, num_chunks, moves, sequence, qstring, n_bases};
    m_cursor.store(0);
    const auto t0 = std::chrono::steady_clock::now();
    {
        std::lock_guard<std::mutex> lock(m_mutex);
        m_done = 0;
        ++m_job_id;
    }
    m_cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lock(m_mutex);
        m_cv_done.wait(lock, [&] { return m_done == (int)m_workers.size(); });
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& w : m_workers) {
        if (!w->error.empty()) {
            const std::string e = w->error;
            w->error.clear();
            throw std::runtime_error("pool worker on device " + std::to_string(w->device) + ": " + e);
        }
    }
    return secs;
}

}  // namespace b200

// Implementation of wave::MultiMatcher (semantics of the reference's
// wave_matching/include/wave/matching/impl/multi_matcher_impl.hpp:19-93).
#ifndef WAVE_MATCHING_IMPL_MULTI_MATCHER_IMPL_HPP
#define WAVE_MATCHING_IMPL_MULTI_MATCHER_IMPL_HPP

namespace wave {

template <class T, class R>
MultiMatcher<T, R>::~MultiMatcher() {
    {
        std::unique_lock<std::mutex> lock(this->ip_mutex);
        this->stop = true;
    }
    this->ip_condition.notify_all();
    this->op_condition.notify_all();
    for (auto &worker : this->pool) worker.join();
    for (T *m : this->matchers) delete m;
}

template <class T, class R>
void MultiMatcher<T, R>::initPool(R params) {
    this->config = params;
    this->matchers.reserve(static_cast<size_t>(this->n_thread));
    // every worker owns its matcher (its own stream / device buffers), as in the reference where
    // each thread gets its own PCL objects
    for (int i = 0; i < this->n_thread; i++) this->matchers.push_back(new T(R(this->config)));
    for (int i = 0; i < this->n_thread; i++) this->pool.emplace_back(&MultiMatcher<T, R>::spin, this, i);
}

template <class T, class R>
void MultiMatcher<T, R>::spin(int threadid) {
    T &matcher = *this->matchers.at(static_cast<size_t>(threadid));
    for (;;) {
        std::tuple<int, PCLPointCloudPtr, PCLPointCloudPtr> job;
        {
            std::unique_lock<std::mutex> lock(this->ip_mutex);
            this->ip_condition.wait(lock, [this] { return this->stop || !this->input.empty(); });
            if (this->stop) return;
            job = this->input.front();
            this->input.pop();
        }
        this->ip_condition.notify_all();  // a slot of the bounded queue is free again
        // the reference's workers never throw (PCL reports failure through hasConverged()); here a
        // device error surfaces as std::runtime_error from the matcher - the job then yields the
        // matcher's previous result/identity instead of ending the process through std::terminate,
        // and remaining_matches still decrements
        try {
            matcher.setRef(std::get<1>(job));
            matcher.setTarget(std::get<2>(job));
            matcher.match();
            matcher.estimateInfo();
        } catch (const std::exception &e) {
            LOG_ERROR("MultiMatcher worker %d: job %d failed: %s", threadid, std::get<0>(job), e.what());
        }
        {
            std::unique_lock<std::mutex> lockop(this->op_mutex);
            this->output.emplace(std::get<0>(job), matcher.getResult(), matcher.getInfo());
            std::unique_lock<std::mutex> lockcnt(this->cnt_mutex);
            --this->remaining_matches;
        }
        this->op_condition.notify_all();
    }
}

template <class T, class R>
void MultiMatcher<T, R>::insert(const int &id, const PCLPointCloudPtr &src, const PCLPointCloudPtr &target) {
    {
        std::unique_lock<std::mutex> lock(this->ip_mutex);
        this->ip_condition.wait(
          lock, [this] { return this->stop || this->input.size() < static_cast<size_t>(this->queue_size); });
        this->input.emplace(id, src, target);
        std::unique_lock<std::mutex> lockcnt(this->cnt_mutex);
        ++this->remaining_matches;
    }
    this->ip_condition.notify_all();
}

template <class T, class R>
bool MultiMatcher<T, R>::done() {
    std::unique_lock<std::mutex> lockcnt(this->cnt_mutex);
    return this->remaining_matches == 0;
}

template <class T, class R>
bool MultiMatcher<T, R>::getResult(int *id, Eigen::Affine3d *transform, Mat6 *info) {
    std::unique_lock<std::mutex> lockop(this->op_mutex);
    this->op_condition.wait(lockop, [this] { return this->stop || !this->output.empty() || this->done(); });
    if (this->output.empty()) return false;
    if (id) *id = std::get<0>(this->output.front());
    if (transform) *transform = std::get<1>(this->output.front());
    if (info) *info = std::get<2>(this->output.front());
    this->output.pop();
    return true;
}

}  // namespace wave

#endif  // WAVE_MATCHING_IMPL_MULTI_MATCHER_IMPL_HPP

// Implementation of wave::MultiMatcher (semantics of the reference's
// wave_matching/include/wave/matching/impl/multi_matcher_impl.hpp:19-93).
#ifndef WAVE_MATCHING_IMPL_MULTI_MATCHER_IMPL_HPP
#define WAVE_MATCHING_IMPL_MULTI_MATCHER_IMPL_HPP

namespace wave {

namespace detail {
// matchers that can read another matcher's target (ICPMatcher::shareTarget)
template <class T, class = void>
struct can_share_target : std::false_type {};
template <class T>
struct can_share_target<T, decltype(void(std::declval<T &>().shareTarget(static_cast<T *>(nullptr))))> : std::true_type {};

template <class T>
void attach_map(T &matcher, const std::vector<T *> &owners) {
    if constexpr (can_share_target<T>::value) {
        for (T *o : owners)
            if (o->device() == matcher.device()) {
                matcher.shareTarget(o);
                return;
            }
    } else {
        (void) matcher;
        (void) owners;
    }
}
}  // namespace detail

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
    for (T *m : this->map_owners) delete m;
}

template <class T, class R>
void MultiMatcher<T, R>::setMap(const PCLPointCloudPtr &map) {
    // one owner per device the workers sit on; each worker then reads its device's owner
    for (T *m : this->map_owners) delete m;
    this->map_owners.clear();
    for (T *worker : this->matchers) {
        T *owner = nullptr;
        for (T *o : this->map_owners)
            if (o->device() == worker->device()) owner = o;
        if (!owner) {
            // a new matcher lands on the next device of the round-robin; keep asking until it is the right one
            for (int tries = 0; tries < 64 && !owner; ++tries) {
                T *cand = new T(R(this->config));
                if (cand->device() == worker->device()) owner = cand;
                else delete cand;
            }
            if (!owner) throw std::runtime_error("MultiMatcher::setMap: no matcher could be placed on the worker's device");
            owner->setTarget(map);
            owner->buildTarget();
            this->map_owners.push_back(owner);
        }
        worker->shareTarget(owner);
    }
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
            if (std::get<2>(job)) matcher.setTarget(std::get<2>(job));
            else detail::attach_map(matcher, this->map_owners);   // null target: the shared map (setMap)
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

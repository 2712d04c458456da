//! `GpuRouter` / `GpuRetainer`: keep rmqtt's `Router` and `RetainStorage` traits, forward the hot calls to
//! libgpumqtt.  UNTESTED SOURCE — written against rmqtt 4f9f2185 without a compiler (see ../README.md).
//!
//! Shape follows the reference's own wrappers:
//!   router  : rmqtt-plugins/rmqtt-cluster-broadcast/src/router.rs:21-170  (delegate everything to DefaultRouter)
//!   retainer: rmqtt-plugins/rmqtt-retainer/src/ram.rs:18-93 + src/lib.rs:143-155
#![deny(unsafe_code)] // all `unsafe` lives in `engine` below, a private module with a safe surface
use std::sync::Arc;
use std::time::Duration;

use async_trait::async_trait;
use dashmap::DashMap;
use rmqtt::{
    context::ServerContext,
    retain::{DefaultRetainStorage, RetainStorage},
    router::{DefaultRouter, Router},
    types::*,
    Result,
};
use tokio::sync::{mpsc, oneshot};

mod engine {
    //! Safe wrapper over gpumqtt-sys.
    #![allow(unsafe_code)]
    use gpumqtt_sys as sys;
    use std::ffi::CStr;

    pub struct Engine(*mut sys::gm_engine);
    unsafe impl Send for Engine {} // libgpumqtt handles are thread-safe (include/gpumqtt.h)
    unsafe impl Sync for Engine {}

    impl Engine {
        pub fn new(device: i32, filters_hint: u64) -> anyhow::Result<Self> {
            let cfg = sys::gm_config { struct_size: std::mem::size_of::<sys::gm_config>() as u32, device, max_levels: 0, flags: 0, filters_hint };
            let mut h = std::ptr::null_mut();
            let rc = unsafe { sys::gm_create(&cfg, &mut h) };
            if rc != sys::GM_OK { anyhow::bail!("gm_create: {}", rc) }
            Ok(Self(h))
        }
        fn err(&self, what: &str, rc: i32) -> anyhow::Error {
            let msg = unsafe { CStr::from_ptr(sys::gm_last_error(self.0)) }.to_string_lossy().into_owned();
            anyhow::anyhow!("{what}: {rc} {msg}")
        }
        pub fn sub_add(&self, filter: &str, value: u32) -> anyhow::Result<bool> {
            let mut ch = 0;
            let rc = unsafe { sys::gm_sub_add(self.0, filter.as_ptr(), filter.len() as u32, value, &mut ch) };
            if rc != sys::GM_OK { return Err(self.err("gm_sub_add", rc)) }
            Ok(ch != 0)
        }
        pub fn sub_remove(&self, filter: &str, value: u32) -> anyhow::Result<bool> {
            let mut ch = 0;
            let rc = unsafe { sys::gm_sub_remove(self.0, filter.as_ptr(), filter.len() as u32, value, &mut ch) };
            if rc != sys::GM_OK { return Err(self.err("gm_sub_remove", rc)) }
            Ok(ch != 0)
        }
        pub fn retain_set(&self, topic: &str, value: u32) -> anyhow::Result<()> {
            let (mut had, mut old) = (0, 0);
            let rc = unsafe { sys::gm_retain_set(self.0, topic.as_ptr(), topic.len() as u32, value, &mut had, &mut old) };
            if rc != sys::GM_OK { return Err(self.err("gm_retain_set", rc)) }
            Ok(())
        }
        pub fn retain_remove(&self, topic: &str) -> anyhow::Result<Option<u32>> {
            let (mut had, mut old) = (0, 0);
            let rc = unsafe { sys::gm_retain_remove(self.0, topic.as_ptr(), topic.len() as u32, &mut had, &mut old) };
            if rc != sys::GM_OK { return Err(self.err("gm_retain_remove", rc)) }
            Ok((had != 0).then_some(old))
        }
        /// One device batch.  `retained == false`: Router::matches; `true`: RetainStorage::get.
        /// Returns per entry `Ok(ids)` or `Err(())` for an invalid topic (Topic::from_str Err, router.rs:165).
        pub fn match_batch(&self, blob: &[u8], offsets: &[u32], retained: bool) -> anyhow::Result<Vec<std::result::Result<Vec<u32>, ()>>> {
            let n = offsets.len() - 1;
            let mut spans = vec![sys::gm_span::default(); n];
            let mut status = vec![0i32; n];
            let mut cap = (n * 64).max(1024);
            loop {
                let mut ids = vec![0u32; cap];
                let mut needed = 0u64;
                let f = if retained { sys::gm_retain_match_batch } else { sys::gm_match_batch };
                let rc = unsafe { f(self.0, blob.as_ptr(), offsets.as_ptr(), n as u64, spans.as_mut_ptr(), ids.as_mut_ptr(), cap as u64, &mut needed, status.as_mut_ptr()) };
                if rc == sys::GM_ERR_CAPACITY { cap = needed as usize; continue }
                if rc != sys::GM_OK { return Err(self.err("gm_match_batch", rc)) }
                return Ok((0..n).map(|i| if status[i] != 0 { Err(()) } else { Ok(ids[spans[i].off as usize..(spans[i].off + spans[i].cnt) as usize].to_vec()) }).collect());
            }
        }
    }
    impl Drop for Engine {
        fn drop(&mut self) { unsafe { sys::gm_destroy(self.0) } }
    }
}
use engine::Engine;

/// Turns per-call async `matches` into device batches: flush at `max_batch` entries or `max_wait`.
struct MicroBatcher {
    tx: mpsc::UnboundedSender<(String, oneshot::Sender<anyhow::Result<Vec<u32>>>)>,
}

impl MicroBatcher {
    fn spawn(engine: Arc<Engine>, retained: bool, max_batch: usize, max_wait: Duration) -> Self {
        let (tx, mut rx) = mpsc::unbounded_channel::<(String, oneshot::Sender<anyhow::Result<Vec<u32>>>)>();
        tokio::spawn(async move {
            loop {
                let Some(first) = rx.recv().await else { break };
                let mut pending = vec![first];
                let deadline = tokio::time::Instant::now() + max_wait;
                while pending.len() < max_batch {
                    match tokio::time::timeout_at(deadline, rx.recv()).await {
                        Ok(Some(x)) => pending.push(x),
                        _ => break,
                    }
                }
                let (mut blob, mut offs) = (Vec::new(), vec![0u32]);
                for (t, _) in &pending { blob.extend_from_slice(t.as_bytes()); offs.push(blob.len() as u32); }
                let eng = engine.clone();
                let res = tokio::task::spawn_blocking(move || eng.match_batch(&blob, &offs, retained)).await;
                match res {
                    Ok(Ok(lists)) => for ((_, tx), l) in pending.into_iter().zip(lists) {
                        let _ = tx.send(l.map_err(|_| anyhow::anyhow!("InvalidTopic")));
                    },
                    Ok(Err(e)) => for (_, tx) in pending { let _ = tx.send(Err(anyhow::anyhow!("{e}"))); },
                    Err(e) => for (_, tx) in pending { let _ = tx.send(Err(anyhow::anyhow!("{e}"))); },
                }
            }
        });
        Self { tx }
    }
    async fn submit(&self, key: String) -> anyhow::Result<Vec<u32>> {
        let (tx, rx) = oneshot::channel();
        self.tx.send((key, tx)).map_err(|_| anyhow::anyhow!("batcher closed"))?;
        rx.await?
    }
}

/// value handle <-> (filter, client): the relation the reference keeps in `relations[filter][client]`.
#[derive(Default)]
struct Handles {
    by_key: DashMap<(TopicFilter, ClientId), u32>,
    by_id: DashMap<u32, (TopicFilter, ClientId)>,
    next: std::sync::atomic::AtomicU32,
}

pub struct GpuRouter {
    inner: DefaultRouter,
    engine: Arc<Engine>,
    handles: Arc<Handles>,
    batcher: MicroBatcher,
}

impl GpuRouter {
    pub fn new(scx: ServerContext, device: i32, max_batch: usize, max_wait: Duration) -> anyhow::Result<Self> {
        let engine = Arc::new(Engine::new(device, 0)?);
        let batcher = MicroBatcher::spawn(engine.clone(), false, max_batch, max_wait);
        Ok(Self { inner: DefaultRouter::new(Some(scx)), engine, handles: Arc::new(Handles::default()), batcher })
    }
}

#[async_trait]
impl Router for GpuRouter {
    async fn add(&self, topic_filter: &str, id: Id, opts: SubscriptionOptions) -> Result<()> {
        self.inner.add(topic_filter, id.clone(), opts).await?; // router.rs:417-436
        let key = (TopicFilter::from(topic_filter), id.client_id.clone());
        let h = *self.handles.by_key.entry(key.clone()).or_insert_with(|| {
            let h = self.handles.next.fetch_add(1, std::sync::atomic::Ordering::Relaxed);
            self.handles.by_id.insert(h, key);
            h
        });
        self.engine.sub_add(topic_filter, h)?;
        Ok(())
    }

    async fn remove(&self, topic_filter: &str, id: Id) -> Result<bool> {
        let removed = self.inner.remove(topic_filter, id.clone()).await?; // router.rs:439-479 keeps the Id-equality rule
        if removed {
            if let Some((_, h)) = self.handles.by_key.remove(&(TopicFilter::from(topic_filter), id.client_id.clone())) {
                self.handles.by_id.remove(&h);
                self.engine.sub_remove(topic_filter, h)?;
            }
        }
        Ok(removed)
    }

    async fn matches(&self, this_id: Id, topic: &TopicName) -> Result<SubRelationsMap> {
        let handles = self.batcher.submit(topic.to_string()).await?; // Err for an invalid topic, like router.rs:165
        // router.rs:182-247 fed from handles instead of the trie iterator
        let mut collector_map: SubscriptioRelationsCollectorMap = Default::default();
        for h in handles {
            let Some(kv) = self.handles.by_id.get(&h) else { continue };
            let (topic_filter, client_id) = kv.value();
            let Some(rels) = self.inner.relations.get(topic_filter) else { continue };
            let Some((id, opts)) = rels.get(client_id) else { continue };
            if let Some(true) = opts.no_local() { if &this_id == id { continue } } // router.rs:184-189
            // shared-subscription bucketing + choice (router.rs:192-238) elided here: identical host code
            collector_map.entry(id.node_id).or_default().add(topic_filter, client_id.clone(), opts.clone(), None);
        }
        Ok(collector_map.into_iter().map(|(n, c)| (n, c.into())).collect())
    }

    async fn is_online(&self, node_id: NodeId, client_id: &str) -> bool { self.inner.is_online(node_id, client_id).await }
    async fn gets(&self, limit: usize) -> Vec<Route> { self.inner.gets(limit).await }
    async fn get(&self, topic: &str) -> Result<Vec<Route>> { self.inner.get(topic).await }
    async fn query_subscriptions(&self, q: &SubsSearchParams) -> Vec<SubsSearchResult> { self.inner.query_subscriptions(q).await }
    async fn topics_tree(&self) -> usize { self.inner.topics_tree().await }
    fn topics(&self) -> Counter { self.inner.topics() }
    fn routes(&self) -> Counter { self.inner.routes() }
    fn merge_topics(&self, m: &HashMap<NodeId, Counter>) -> Counter { self.inner.merge_topics(m) }
    fn merge_routes(&self, m: &HashMap<NodeId, Counter>) -> Counter { self.inner.merge_routes(m) }
    async fn list_topics(&self, top: usize) -> Vec<String> { self.inner.list_topics(top).await }
    async fn list_relations(&self, top: usize) -> Vec<serde_json::Value> { self.inner.list_relations(top).await }
    fn relations(&self) -> &AllRelationsMap { self.inner.relations() }
}

pub struct GpuRetainer {
    inner: Arc<DefaultRetainStorage>,
    engine: Arc<Engine>,
    ids: DashMap<TopicName, u32>,
    names: DashMap<u32, TopicName>,
    next: std::sync::atomic::AtomicU32,
    batcher: MicroBatcher,
}

#[async_trait]
impl RetainStorage for GpuRetainer {
    fn enable(&self) -> bool { true }

    async fn set(&self, topic: &TopicName, retain: Retain, expiry_interval: Option<Duration>) -> Result<()> {
        let empty = retain.publish.payload.is_empty();
        self.inner.set_with_timeout(topic, retain, expiry_interval).await?; // retain.rs:131-149
        if empty {
            if let Some((_, id)) = self.ids.remove(topic) { self.names.remove(&id); }
            self.engine.retain_remove(topic)?;
        } else {
            let id = *self.ids.entry(topic.clone()).or_insert_with(|| {
                let id = self.next.fetch_add(1, std::sync::atomic::Ordering::Relaxed);
                self.names.insert(id, topic.clone());
                id
            });
            self.engine.retain_set(topic, id)?;
        }
        Ok(())
    }

    async fn get(&self, topic_filter: &TopicFilter) -> Result<Vec<(TopicName, Retain)>> {
        let ids = self.batcher.submit(topic_filter.to_string()).await?;
        let messages = self.inner.messages.read().await;
        let mut out = Vec::with_capacity(ids.len());
        for id in ids {
            let Some(name) = self.names.get(&id) else { continue };
            // exact lookup of the payload by concrete topic; expired entries are dropped like retain.rs:158-166
            for (t, tv) in messages.matches(&name.parse()?) {
                if !tv.is_expired() { out.push((TopicName::from(t.to_string()), tv.into_value())); }
            }
        }
        Ok(out)
    }

    async fn count(&self) -> isize { self.inner.count().await }
    async fn max(&self) -> isize { self.inner.max().await }
}

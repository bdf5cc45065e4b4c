//! UNVERIFIED SOURCE — never compiled (no rustc in the build image).
//!
//! Thin Rust host layer keeping the hnsw_rs API surface (`Hnsw<f32, D>`, `AnnT`, `FilterT`, `Neighbour`)
//! on top of the C ABI of libhnsw_b200.so (include/hnsw_b200.h).  Names, argument order and meaning follow
//! hnsw_rs 0.3.4: `src/hnsw.rs:771-777,1069-1071,1224-1238,1487-1635`, `src/api.rs:13-38`, `src/filter.rs:7-24`.
#![allow(non_camel_case_types)]
use std::marker::PhantomData;
use std::os::raw::{c_int, c_void};

pub type DataId = usize;

#[repr(C)]
pub struct HnswApif32 {
    _private: [u8; 0],
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct Neighbour_api {
    pub id: usize,
    pub d: f32,
}
#[repr(C)]
pub struct Neighbourhood_api {
    pub nbgh: i64,
    pub neighbours: *const Neighbour_api,
}
#[repr(C)]
pub struct Vec_api<T> {
    pub len: i64,
    pub ptr: *const T,
}

extern "C" {
    fn new_hnsw_f32(max_nb_conn: usize, ef_const: usize, namelen: usize, cdistname: *const u8, max_elements: usize,
                    max_layer: usize) -> *const HnswApif32;
    fn drop_hnsw_f32(p: *const HnswApif32);
    fn insert_f32(h: *mut HnswApif32, len: usize, data: *const f32, id: usize);
    fn parallel_insert_f32(h: *mut HnswApif32, nb_vec: usize, vec_len: usize, datas: *mut *const f32, ids: *const usize);
    fn parallel_search_neighbours_f32(h: *const HnswApif32, nb_vec: usize, vec_len: i64, data: *mut *const f32,
                                      knbn: usize, ef_search: usize) -> *const Vec_api<Neighbourhood_api>;
    fn hnsw_b200_free_vec_api(p: *const Vec_api<Neighbourhood_api>);
    fn hnsw_b200_search_flat(h: *const HnswApif32, queries: *const f32, nq: u64, dim: u64, knbn: u64, ef: u64,
                             filter_mode: c_int, filter_ids: *const u64, nfilter: u64,
                             f: Option<extern "C" fn(u64, *mut c_void) -> c_int>, ctx: *mut c_void, out_ids: *mut u64,
                             out_dist: *mut f32, out_internal: *mut u32, out_pid: *mut i32, out_counts: *mut i32) -> c_int;
    fn hnsw_b200_set_extend_candidates(h: *mut HnswApif32, flag: c_int) -> c_int;
    fn hnsw_b200_set_keeping_pruned(h: *mut HnswApif32, flag: c_int) -> c_int;
    fn hnsw_b200_modify_level_scale(h: *mut HnswApif32, scale: f64) -> c_int;
    fn hnsw_b200_set_searching_mode(h: *mut HnswApif32, flag: c_int) -> c_int;
    fn hnsw_b200_get_nb_point(h: *const HnswApif32) -> u64;
    // multi-GPU (include/hnsw_b200.h "Multi-GPU search"): one process drives several devices
    fn hnsw_b200_replicate(h: *mut HnswApif32, ndev: c_int, devices: *const c_int) -> c_int;
    fn hnsw_b200_replica_count(h: *const HnswApif32) -> c_int;
}

/// hnsw.rs:46
#[derive(Debug, Clone, Copy, Default, PartialEq, Eq)]
pub struct PointId(pub u8, pub i32);

/// hnsw.rs:98-107
#[derive(Debug, Clone, Copy, Default)]
pub struct Neighbour {
    pub d_id: DataId,
    pub distance: f32,
    pub p_id: PointId,
}

/// filter.rs:7-9
pub trait FilterT {
    fn hnsw_filter(&self, id: &DataId) -> bool;
}
impl FilterT for Vec<usize> {
    fn hnsw_filter(&self, id: &DataId) -> bool {
        self.binary_search(id).is_ok()
    }
}
impl<F: Fn(&DataId) -> bool> FilterT for F {
    fn hnsw_filter(&self, id: &DataId) -> bool {
        self(id)
    }
}

/// Distance marker types: the kernels are selected by NAME, as in libext.rs:468-520.
pub trait DistName {
    const NAME: &'static str;
}
macro_rules! dist { ($t:ident) => { #[derive(Default, Clone, Copy)] pub struct $t; impl DistName for $t { const NAME: &'static str = stringify!($t); } } }
dist!(DistL1); dist!(DistL2); dist!(DistDot); dist!(DistCosine); dist!(DistHellinger); dist!(DistJeffreys); dist!(DistJensenShannon);

pub struct Hnsw<D: DistName> {
    h: *mut HnswApif32,
    _d: PhantomData<D>,
}
unsafe impl<D: DistName> Send for Hnsw<D> {}
unsafe impl<D: DistName> Sync for Hnsw<D> {}

extern "C" fn filter_trampoline(id: u64, ctx: *mut c_void) -> c_int {
    let f: &&dyn FilterT = unsafe { &*(ctx as *const &dyn FilterT) };
    f.hnsw_filter(&(id as usize)) as c_int
}

impl<D: DistName> Hnsw<D> {
    /// Hnsw::new, hnsw.rs:771-777
    pub fn new(max_nb_connection: usize, max_elements: usize, max_layer: usize, ef_construction: usize, _f: D) -> Self {
        let name = D::NAME.as_bytes();
        let h = unsafe { new_hnsw_f32(max_nb_connection, ef_construction, name.len(), name.as_ptr(), max_elements, max_layer) };
        assert!(!h.is_null(), "libhnsw_b200: no usable CUDA device or bad parameters (there is no CPU fallback)");
        Hnsw { h: h as *mut HnswApif32, _d: PhantomData }
    }
    pub fn get_nb_point(&self) -> usize { unsafe { hnsw_b200_get_nb_point(self.h) as usize } }
    pub fn set_extend_candidates(&mut self, flag: bool) { unsafe { hnsw_b200_set_extend_candidates(self.h, flag as c_int); } }
    pub fn set_keeping_pruned(&mut self, flag: bool) { unsafe { hnsw_b200_set_keeping_pruned(self.h, flag as c_int); } }
    pub fn modify_level_scale(&mut self, s: f64) { unsafe { hnsw_b200_modify_level_scale(self.h, s); } }
    pub fn set_searching_mode(&mut self, flag: bool) { unsafe { hnsw_b200_set_searching_mode(self.h, flag as c_int); } }
    /// Extension: copy the index to `devices[1..]` (devices[0] = the device it lives on); `parallel_search` then shards its
    /// batch over all of them, one call as on the CPU (hnsw.rs:1612-1635).
    pub fn replicate(&mut self, devices: &[i32]) -> Result<(), i32> {
        let r = unsafe { hnsw_b200_replicate(self.h, devices.len() as c_int, devices.as_ptr()) };
        if r == 0 { Ok(()) } else { Err(r) }
    }
    pub fn replica_count(&self) -> usize { unsafe { hnsw_b200_replica_count(self.h) as usize } }

    /// hnsw.rs:1069-1071
    pub fn insert(&self, datav_with_id: (&[f32], usize)) {
        unsafe { insert_f32(self.h, datav_with_id.0.len(), datav_with_id.0.as_ptr(), datav_with_id.1) }
    }
    /// hnsw.rs:1224-1230
    pub fn parallel_insert(&self, datas: &[(&Vec<f32>, usize)]) {
        if datas.is_empty() { return; }
        let mut ptrs: Vec<*const f32> = datas.iter().map(|d| d.0.as_ptr()).collect();
        let ids: Vec<usize> = datas.iter().map(|d| d.1).collect();
        unsafe { parallel_insert_f32(self.h, datas.len(), datas[0].0.len(), ptrs.as_mut_ptr(), ids.as_ptr()) }
    }
    /// hnsw.rs:1597-1599
    pub fn search(&self, data: &[f32], knbn: usize, ef_arg: usize) -> Vec<Neighbour> {
        self.search_filter(data, knbn, ef_arg, None)
    }
    /// hnsw.rs:1487-1580.  Closures are evaluated once per stored origin id by the library (device bitmap).
    pub fn search_filter(&self, data: &[f32], knbn: usize, ef_arg: usize, filter: Option<&dyn FilterT>) -> Vec<Neighbour> {
        let mut ids = vec![0u64; knbn];
        let mut ds = vec![0f32; knbn];
        let mut pid = vec![0i32; 2 * knbn];
        let mut cnt = 0i32;
        let (mode, cb, ctx) = match filter.as_ref() {
            None => (0, None, std::ptr::null_mut()),
            Some(f) => (2, Some(filter_trampoline as extern "C" fn(u64, *mut c_void) -> c_int), f as *const &dyn FilterT as *mut c_void),
        };
        let r = unsafe {
            hnsw_b200_search_flat(self.h, data.as_ptr(), 1, data.len() as u64, knbn as u64, ef_arg as u64, mode,
                                  std::ptr::null(), 0, cb, ctx, ids.as_mut_ptr(), ds.as_mut_ptr(), std::ptr::null_mut(),
                                  pid.as_mut_ptr(), &mut cnt)
        };
        assert_eq!(r, 0, "hnsw_b200_search_flat failed");
        (0..cnt as usize).map(|j| Neighbour { d_id: ids[j] as usize, distance: ds[j], p_id: PointId(pid[2 * j] as u8, pid[2 * j + 1]) }).collect()
    }
    /// hnsw.rs:1612-1635: one answer per request, in input order
    pub fn parallel_search(&self, datas: &[Vec<f32>], knbn: usize, ef: usize) -> Vec<Vec<Neighbour>> {
        if datas.is_empty() { return Vec::new(); }
        let mut ptrs: Vec<*const f32> = datas.iter().map(|d| d.as_ptr()).collect();
        let res = unsafe { parallel_search_neighbours_f32(self.h, datas.len(), datas[0].len() as i64, ptrs.as_mut_ptr(), knbn, ef) };
        assert!(!res.is_null());
        let v = unsafe { &*res };
        let hoods = unsafe { std::slice::from_raw_parts(v.ptr, v.len as usize) };
        let out = hoods.iter().map(|h| {
            let nb = unsafe { std::slice::from_raw_parts(h.neighbours, h.nbgh as usize) };
            nb.iter().map(|n| Neighbour { d_id: n.id, distance: n.d, p_id: PointId::default() }).collect()
        }).collect();
        unsafe { hnsw_b200_free_vec_api(res) };
        out
    }
}

impl<D: DistName> Drop for Hnsw<D> {
    fn drop(&mut self) { unsafe { drop_hnsw_f32(self.h) } }
}

/// api.rs:13-38
pub trait AnnT {
    type Val;
    fn insert_data(&mut self, data: &[Self::Val], id: usize);
    fn search_neighbours(&self, data: &[Self::Val], knbn: usize, ef_s: usize) -> Vec<Neighbour>;
    fn parallel_insert_data(&mut self, data: &[(&Vec<Self::Val>, usize)]);
    fn parallel_search_neighbours(&self, data: &[Vec<Self::Val>], knbn: usize, ef_s: usize) -> Vec<Vec<Neighbour>>;
}
impl<D: DistName> AnnT for Hnsw<D> {
    type Val = f32;
    fn insert_data(&mut self, data: &[f32], id: usize) { self.insert((data, id)) }
    fn search_neighbours(&self, data: &[f32], knbn: usize, ef_s: usize) -> Vec<Neighbour> { self.search(data, knbn, ef_s) }
    fn parallel_insert_data(&mut self, data: &[(&Vec<f32>, usize)]) { self.parallel_insert(data) }
    fn parallel_search_neighbours(&self, data: &[Vec<f32>], knbn: usize, ef_s: usize) -> Vec<Vec<Neighbour>> { self.parallel_search(data, knbn, ef_s) }
}

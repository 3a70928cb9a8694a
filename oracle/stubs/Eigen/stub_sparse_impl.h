// oracle/stubs/Eigen/stub_sparse_impl.h — TEST INFRASTRUCTURE ONLY, NOT Eigen (see Eigen/Core here).
// Sparse matrices, sparse vectors and the tiny dense matrix the reference's design code touches.
#ifndef ORACLE_STUB_EIGEN_SPARSE_IMPL
#define ORACLE_STUB_EIGEN_SPARSE_IMPL

namespace Eigen {

// Sparse matrix: outer_[k] is the sorted (inner index, value) list of outer slice k
// (a column when O == ColMajor, a row when O == RowMajor).
template <class S, int O = ColMajor, class I = int> class SparseMatrix {
 public:
  typedef std::vector<std::pair<Index, S> > Slice;
  SparseMatrix() : rows_(0), cols_(0) {}
  SparseMatrix(Index r, Index c) : rows_(r), cols_(c), outer_(O == RowMajor ? r : c) {}
  template <int O2, class I2> SparseMatrix(const SparseMatrix<S, O2, I2>& m) : rows_(m.rows()), cols_(m.cols()) {
    outer_.assign(O == RowMajor ? rows_ : cols_, Slice());
    for (Index k = 0; k < m.outerSize(); ++k)
      for (typename SparseMatrix<S, O2, I2>::InnerIterator it(m, k); it; ++it)
        outer_[O == RowMajor ? it.row() : it.col()].push_back(std::make_pair(O == RowMajor ? it.col() : it.row(), it.value()));
    sort_slices();
  }

  Index rows() const { return rows_; }
  Index cols() const { return cols_; }
  Index outerSize() const { return (Index)outer_.size(); }
  Index innerSize() const { return O == RowMajor ? cols_ : rows_; }
  Index nonZeros() const { Index n = 0; for (const Slice& s : outer_) n += (Index)s.size(); return n; }
  void makeCompressed() {}
  void setIdentity() {
    for (Slice& s : outer_) s.clear();
    for (Index i = 0; i < std::min(rows_, cols_); ++i) outer_[i].push_back(std::make_pair(i, S(1)));
  }
  template <class It> void setFromTriplets(It b, It e) {
    for (Slice& s : outer_) s.clear();
    for (It t = b; t != e; ++t) {
      const Index r = t->row(), c = t->col();
      if (r < 0 || r >= rows_ || c < 0 || c >= cols_) throw std::out_of_range("stub Eigen: triplet out of range");
      outer_[O == RowMajor ? r : c].push_back(std::make_pair(O == RowMajor ? c : r, t->value()));
    }
    sort_slices();
    for (Slice& s : outer_) {            // duplicates are summed, as setFromTriplets does
      Slice u;
      for (const auto& e2 : s) { if (!u.empty() && u.back().first == e2.first) u.back().second += e2.second; else u.push_back(e2); }
      s.swap(u);
    }
  }

  class InnerIterator {
   public:
    InnerIterator(const SparseMatrix& m, Index k) : s_(&m.outer_[k]), k_(k), i_(0) {}
    InnerIterator& operator++() { ++i_; return *this; }
    operator bool() const { return i_ < s_->size(); }
    Index index() const { return (*s_)[i_].first; }
    Index row() const { return O == RowMajor ? k_ : index(); }
    Index col() const { return O == RowMajor ? index() : k_; }
    S value() const { return (*s_)[i_].second; }
   private:
    const Slice* s_; Index k_; std::size_t i_;
  };

  S coeff(Index r, Index c) const {
    const Slice& s = outer_[O == RowMajor ? r : c];
    const Index in = O == RowMajor ? c : r;
    for (const auto& e : s) if (e.first == in) return e.second;
    return S(0);
  }
  SparseMatrix<S, ColMajor, int> col(Index j) const {
    SparseMatrix<S, ColMajor, int> r(rows_, 1);
    std::vector<Triplet<S> > t;
    for (Index k = 0; k < outerSize(); ++k)
      for (InnerIterator it(*this, k); it; ++it) if (it.col() == j) t.push_back(Triplet<S>(it.row(), 0, it.value()));
    r.setFromTriplets(t.begin(), t.end());
    return r;
  }
  // transposing flips the storage order and keeps the slices
  SparseMatrix<S, 1 - O, I> transpose() const {
    SparseMatrix<S, 1 - O, I> r(cols_, rows_);
    r.raw() = outer_;
    return r;
  }
  Arr<S> diagonal() const {
    Arr<S> d = Arr<S>::Zero(std::min(rows_, cols_));
    for (Index i = 0; i < d.size(); ++i) d(i) = coeff(i, i);
    return d;
  }
  std::vector<Slice>& raw() { return outer_; }
  const std::vector<Slice>& raw() const { return outer_; }
  // row-wise copy (sorted by column), used by the products
  std::vector<Slice> by_rows() const {
    if (O == RowMajor) return outer_;
    std::vector<Slice> r(rows_);
    for (Index k = 0; k < outerSize(); ++k) for (const auto& e : outer_[k]) r[e.first].push_back(std::make_pair(k, e.second));
    return r;
  }

 private:
  void sort_slices() {
    for (Slice& s : outer_)
      std::stable_sort(s.begin(), s.end(), [](const std::pair<Index, S>& a, const std::pair<Index, S>& b) { return a.first < b.first; });
  }
  Index rows_, cols_;
  std::vector<Slice> outer_;
};

// sparse * sparse: C(i,j) = sum_k A(i,k) B(k,j), k increasing, accumulated from 0
template <class S, int O1, class I1, int O2, class I2>
SparseMatrix<S, ColMajor, int> operator*(const SparseMatrix<S, O1, I1>& a, const SparseMatrix<S, O2, I2>& b) {
  if (a.cols() != b.rows()) throw std::invalid_argument("stub Eigen: sparse product size mismatch");
  const auto ar = a.by_rows(), br = b.by_rows();
  std::vector<Triplet<S> > t;
  for (Index i = 0; i < a.rows(); ++i) {
    std::map<Index, S> acc;
    for (const auto& ea : ar[i]) for (const auto& eb : br[ea.first]) {
      auto it = acc.find(eb.first);
      if (it == acc.end()) acc[eb.first] = S(0) + ea.second * eb.second; else it->second = it->second + ea.second * eb.second;
    }
    for (const auto& kv : acc) t.push_back(Triplet<S>(i, kv.first, kv.second));
  }
  SparseMatrix<S, ColMajor, int> c(a.rows(), b.cols());
  c.setFromTriplets(t.begin(), t.end());
  return c;
}
// sparse * diagonal scales the columns, diagonal * sparse the rows
template <class S, int O, class I> SparseMatrix<S, O, I> operator*(const SparseMatrix<S, O, I>& a, const Diag<S>& d) {
  SparseMatrix<S, O, I> r(a);
  for (Index k = 0; k < r.outerSize(); ++k) for (auto& e : r.raw()[k]) e.second = e.second * d.d(O == RowMajor ? e.first : k);
  return r;
}
template <class S, int O, class I> SparseMatrix<S, O, I> operator*(const Diag<S>& d, const SparseMatrix<S, O, I>& a) {
  SparseMatrix<S, O, I> r(a);
  for (Index k = 0; k < r.outerSize(); ++k) for (auto& e : r.raw()[k]) e.second = d.d(O == RowMajor ? k : e.first) * e.second;
  return r;
}
// sparse * dense vector: y(i) = sum_j A(i,j) x(j), j increasing, accumulated from 0
template <class S, int O, class I> Arr<S> operator*(const SparseMatrix<S, O, I>& a, const Arr<S>& x) {
  if (a.cols() != x.size()) throw std::invalid_argument("stub Eigen: sparse*vector size mismatch");
  Arr<S> y = Arr<S>::Zero(a.rows());
  const auto ar = a.by_rows();
  for (Index i = 0; i < a.rows(); ++i) for (const auto& e : ar[i]) y(i) = y(i) + e.second * x(e.first);
  return y;
}
template <class S, int O, class I> std::ostream& operator<<(std::ostream& os, const SparseMatrix<S, O, I>& m) {
  for (Index i = 0; i < m.rows(); ++i) { for (Index j = 0; j < m.cols(); ++j) os << m.coeff(i, j) << " "; os << "\n"; }
  return os;
}

template <class T> template <int O, class I> Arr<T>::Arr(const SparseMatrix<T, O, I>& m) {
  if (m.cols() != 1) throw std::invalid_argument("stub Eigen: dense vector from a sparse matrix needs one column");
  *this = Arr<T>::Zero(m.rows());
  for (Index k = 0; k < m.outerSize(); ++k) for (typename SparseMatrix<T, O, I>::InnerIterator it(m, k); it; ++it) (*this)(it.row()) = it.value();
}
template <class T> SparseMatrix<T, ColMajor, int> Arr<T>::sparseView() const {
  SparseMatrix<T, ColMajor, int> r(size(), 1);
  std::vector<Triplet<T> > t;
  for (Index i = 0; i < size(); ++i) if (v_[i] != T(0)) t.push_back(Triplet<T>(i, 0, v_[i]));
  r.setFromTriplets(t.begin(), t.end());
  return r;
}

// sparse vector built from a one-column (or one-row) sparse matrix
template <class S> class SparseVector {
 public:
  template <int O, class I> SparseVector(const SparseMatrix<S, O, I>& m) {
    const bool colvec = m.cols() == 1;
    if (!colvec && m.rows() != 1) throw std::invalid_argument("stub Eigen: SparseVector needs a vector-shaped matrix");
    n_ = colvec ? m.rows() : m.cols();
    for (Index k = 0; k < m.outerSize(); ++k)
      for (typename SparseMatrix<S, O, I>::InnerIterator it(m, k); it; ++it) e_.push_back(std::make_pair(colvec ? it.row() : it.col(), it.value()));
    std::sort(e_.begin(), e_.end(), [](const std::pair<Index, S>& a, const std::pair<Index, S>& b) { return a.first < b.first; });
  }
  Index size() const { return n_; }
  Index rows() const { return n_; }
  S& coeffRef(Index i) {
    for (auto& e : e_) if (e.first == i) return e.second;
    e_.push_back(std::make_pair(i, S(0)));
    std::sort(e_.begin(), e_.end(), [](const std::pair<Index, S>& a, const std::pair<Index, S>& b) { return a.first < b.first; });
    return coeffRef(i);
  }
  class InnerIterator {
   public:
    explicit InnerIterator(const SparseVector& v) : v_(&v), i_(0) {}
    InnerIterator& operator++() { ++i_; return *this; }
    operator bool() const { return i_ < v_->e_.size(); }
    Index index() const { return v_->e_[i_].first; }
    S value() const { return v_->e_[i_].second; }
   private:
    const SparseVector* v_; std::size_t i_;
  };
 private:
  Index n_;
  std::vector<std::pair<Index, S> > e_;
};

// ---------------------------------------------------------------------------------------------
// dense matrix (column-major), only what build_difference_design and the debug prints use
class MatrixXd {
 public:
  MatrixXd() : r_(0), c_(0) {}
  MatrixXd(Index r, Index c) : r_(r), c_(c), v_(r * c, 0.0) {}
  template <int O, class I> MatrixXd(const SparseMatrix<double, O, I>& m) : r_(m.rows()), c_(m.cols()), v_(m.rows() * m.cols(), 0.0) {
    for (Index k = 0; k < m.outerSize(); ++k)
      for (typename SparseMatrix<double, O, I>::InnerIterator it(m, k); it; ++it) (*this)(it.row(), it.col()) = it.value();
  }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  // always range-checked: build_difference_design writes block(0,1,D,ref-1), which is out of range when
  // ref exceeds the number of conditions (undefined behaviour with real Eigen under NDEBUG); here it throws
  double& operator()(Index i, Index j) { check(i, j); return v_[j * r_ + i]; }
  double operator()(Index i, Index j) const { check(i, j); return v_[j * r_ + i]; }
  void check(Index i, Index j) const { if (i < 0 || i >= r_ || j < 0 || j >= c_) throw std::out_of_range("stub Eigen: dense matrix index out of range"); }
  const MatrixXd& eval() const { return *this; }
  MatrixXd leftCols(Index k) const { return block_copy(0, 0, r_, k); }
  MatrixXd topRows(Index k) const { return block_copy(0, 0, k, c_); }
  MatrixXd block_copy(Index i0, Index j0, Index nr, Index nc) const {
    MatrixXd b(nr, nc);
    for (Index j = 0; j < nc; ++j) for (Index i = 0; i < nr; ++i) b(i, j) = (*this)(i0 + i, j0 + j);
    return b;
  }
  class BlockRef {
   public:
    BlockRef(MatrixXd& m, Index i0, Index j0, Index nr, Index nc) : m_(m), i0_(i0), j0_(j0), nr_(nr), nc_(nc) {}
    BlockRef& operator=(const MatrixXd& s) {
      if (s.rows() != nr_ || s.cols() != nc_) throw std::invalid_argument("stub Eigen: block size mismatch");
      for (Index j = 0; j < nc_; ++j) for (Index i = 0; i < nr_; ++i) m_(i0_ + i, j0_ + j) = s(i, j);
      return *this;
    }
    BlockRef& operator=(const Arr<double>& s) {
      if (nc_ != 1 || s.size() != nr_) throw std::invalid_argument("stub Eigen: column size mismatch");
      for (Index i = 0; i < nr_; ++i) m_(i0_ + i, j0_) = s(i);
      return *this;
    }
   private:
    MatrixXd& m_; Index i0_, j0_, nr_, nc_;
  };
  BlockRef block(Index i0, Index j0, Index nr, Index nc) { return BlockRef(*this, i0, j0, nr, nc); }
  BlockRef col(Index j) { return BlockRef(*this, 0, j, r_, 1); }
  SparseMatrix<double, ColMajor, int> sparseView() const {
    SparseMatrix<double, ColMajor, int> s(r_, c_);
    std::vector<Triplet<double> > t;
    for (Index j = 0; j < c_; ++j) for (Index i = 0; i < r_; ++i) if ((*this)(i, j) != 0.0) t.push_back(Triplet<double>(i, j, (*this)(i, j)));
    s.setFromTriplets(t.begin(), t.end());
    return s;
  }
  // comma initialiser by whole columns (only used inside the reference's disabled debug prints)
  class Comma {
   public:
    Comma(MatrixXd& m) : m_(m), j_(0) {}
    Comma& operator,(const Arr<double>& a) { for (Index i = 0; i < m_.rows(); ++i) m_(i, j_) = a(i); ++j_; return *this; }
    MatrixXd finished() const { return m_; }
   private:
    MatrixXd& m_; Index j_;
  };
  Comma operator<<(const Arr<double>& a) { Comma c(*this); c, a; return c; }
 private:
  Index r_, c_;
  std::vector<double> v_;
};
inline std::ostream& operator<<(std::ostream& os, const MatrixXd& m) {
  for (Index i = 0; i < m.rows(); ++i) { for (Index j = 0; j < m.cols(); ++j) os << m(i, j) << " "; os << "\n"; }
  return os;
}

}  // namespace Eigen
#endif

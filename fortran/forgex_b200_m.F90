! forgex_b200_m -- ISO_C_BINDING interface to libforgex_b200.so (include/forgex_b200.h) and the batch forms
! that Forgex's Fortran API gains on top of it.
!
! STATUS: UNVERIFIED.  This image has no Fortran front end (no gfortran / flang / nvfortran / ifx), here or on
! the GPU box, so this file has never been compiled.  It is kept deliberately thin: every procedure is one
! interface block or a few lines of marshalling around a C entry point that IS exercised by the Python and C
! tests.  INTEGRATION.md shows where it plugs into the reference (src/forgex.F90, src/api_internal_m.F90).
!
! What stays as it is in Forgex: operator(.in.), operator(.match.), regex, regex_f, is_valid_regex
! (src/forgex.F90:24-54).  What is new: a compiled-pattern handle and batch generics
!     fx_pattern_t            compile once (the reference recompiles per call: src/forgex.F90:98, :139-140)
!     match_batch / in_batch  one pattern against n fixed-length strings (character(len=*) :: strs(:)) or against
!                             a flat buffer + offsets (variable length; trailing blanks are text to Forgex)
!     regex_batch             (from, to) per string
!     regex_buffer            one pattern against one huge buffer, 64-bit positions
module forgex_b200_m
   use, intrinsic :: iso_c_binding
   use, intrinsic :: iso_fortran_env, only: int64
   implicit none
   private

   integer(c_int), parameter, public :: FX_OP_MATCH = 0, FX_OP_IN = 1, FX_OP_REGEX = 2
   integer(c_int), parameter, public :: FX_OK = 0
   integer(c_int), parameter, public :: FX_ERR_TREE_NODE_LIMIT = 101, FX_ERR_DFA_STATE_CAP = 102, &
                                        FX_ERR_PREFILTER_UNSUPPORTED = 103, FX_ERR_BAD_ARGUMENT = 104, &
                                        FX_ERR_NO_DEVICE = 105

   type, public :: fx_pattern_t
      type(c_ptr) :: handle = c_null_ptr
      integer(c_int) :: status = -1
   contains
      procedure :: compile => pattern__compile
      procedure :: free    => pattern__free
   end type fx_pattern_t

   public :: match_batch, in_batch, regex_batch, regex_buffer

   interface match_batch
      module procedure :: match_batch__fixed, match_batch__ragged
   end interface
   interface in_batch
      module procedure :: in_batch__fixed, in_batch__ragged
   end interface

   ! ---- C entry points (include/forgex_b200.h) -------------------------------------------------------
   interface
      function fx_compile(pattern, plen, op, out) bind(c, name='fx_compile') result(status)
         import :: c_char, c_int64_t, c_int, c_ptr
         character(kind=c_char), intent(in) :: pattern(*)
         integer(c_int64_t), value :: plen
         integer(c_int), value :: op
         type(c_ptr), intent(out) :: out
         integer(c_int) :: status
      end function
      function fx_pattern_free(p) bind(c, name='fx_pattern_free') result(status)
         import :: c_ptr, c_int
         type(c_ptr), value :: p
         integer(c_int) :: status
      end function
      function fx_match_fixed(p, buf, n, stride, out) bind(c, name='fx_match_fixed') result(status)
         import :: c_ptr, c_char, c_int64_t, c_int8_t, c_int
         type(c_ptr), value :: p
         character(kind=c_char), intent(in) :: buf(*)
         integer(c_int64_t), value :: n, stride
         integer(c_int8_t), intent(out) :: out(*)
         integer(c_int) :: status
      end function
      function fx_in_fixed(p, buf, n, stride, out) bind(c, name='fx_in_fixed') result(status)
         import :: c_ptr, c_char, c_int64_t, c_int8_t, c_int
         type(c_ptr), value :: p
         character(kind=c_char), intent(in) :: buf(*)
         integer(c_int64_t), value :: n, stride
         integer(c_int8_t), intent(out) :: out(*)
         integer(c_int) :: status
      end function
      function fx_match_batch(p, buf, offsets, n, out) bind(c, name='fx_match_batch') result(status)
         import :: c_ptr, c_char, c_int64_t, c_int8_t, c_int
         type(c_ptr), value :: p
         character(kind=c_char), intent(in) :: buf(*)
         integer(c_int64_t), intent(in) :: offsets(*)
         integer(c_int64_t), value :: n
         integer(c_int8_t), intent(out) :: out(*)
         integer(c_int) :: status
      end function
      function fx_in_batch(p, buf, offsets, n, out) bind(c, name='fx_in_batch') result(status)
         import :: c_ptr, c_char, c_int64_t, c_int8_t, c_int
         type(c_ptr), value :: p
         character(kind=c_char), intent(in) :: buf(*)
         integer(c_int64_t), intent(in) :: offsets(*)
         integer(c_int64_t), value :: n
         integer(c_int8_t), intent(out) :: out(*)
         integer(c_int) :: status
      end function
      function fx_regex_batch(p, buf, offsets, n, from, to) bind(c, name='fx_regex_batch') result(status)
         import :: c_ptr, c_char, c_int64_t, c_int
         type(c_ptr), value :: p
         character(kind=c_char), intent(in) :: buf(*)
         integer(c_int64_t), intent(in) :: offsets(*)
         integer(c_int64_t), value :: n
         integer(c_int64_t), intent(out) :: from(*), to(*)
         integer(c_int) :: status
      end function
      function fx_regex_buffer(p, buf, length, from, to) bind(c, name='fx_regex_buffer') result(status)
         import :: c_ptr, c_char, c_int64_t, c_int
         type(c_ptr), value :: p
         character(kind=c_char), intent(in) :: buf(*)
         integer(c_int64_t), value :: length
         integer(c_int64_t), intent(out) :: from, to
         integer(c_int) :: status
      end function
      ! one pattern, one text: drop-in bodies for operator__in / operator__match / subroutine__regex
      function fx_in(pattern, plen, text, tlen, res) bind(c, name='fx_in') result(status)
         import :: c_char, c_int64_t, c_int
         character(kind=c_char), intent(in) :: pattern(*), text(*)
         integer(c_int64_t), value :: plen, tlen
         integer(c_int), intent(out) :: res
         integer(c_int) :: status
      end function
      function fx_match(pattern, plen, text, tlen, res) bind(c, name='fx_match') result(status)
         import :: c_char, c_int64_t, c_int
         character(kind=c_char), intent(in) :: pattern(*), text(*)
         integer(c_int64_t), value :: plen, tlen
         integer(c_int), intent(out) :: res
         integer(c_int) :: status
      end function
      function fx_regex(pattern, plen, text, tlen, from, to, length, syntax_status) bind(c, name='fx_regex') result(status)
         import :: c_char, c_int64_t, c_int
         character(kind=c_char), intent(in) :: pattern(*), text(*)
         integer(c_int64_t), value :: plen, tlen
         integer(c_int64_t), intent(out) :: from, to, length
         integer(c_int), intent(out) :: syntax_status
         integer(c_int) :: status
      end function
   end interface
   public :: fx_in, fx_match, fx_regex

contains

   subroutine pattern__compile(self, pattern, op)
      class(fx_pattern_t), intent(inout) :: self
      character(*), intent(in) :: pattern      ! NOT trimmed here: each entry point applies Forgex's own
      integer(c_int), intent(in) :: op          ! preprocessing (src/forgex.F90:95, :182-190, :260) inside fx_compile
      self%status = fx_compile(pattern, int(len(pattern), c_int64_t), op, self%handle)
   end subroutine pattern__compile

   subroutine pattern__free(self)
      class(fx_pattern_t), intent(inout) :: self
      integer(c_int) :: ignore
      if (c_associated(self%handle)) ignore = fx_pattern_free(self%handle)
      self%handle = c_null_ptr
   end subroutine pattern__free

   ! fixed-length strings: a Fortran character array is already one contiguous buffer of n*len bytes
   subroutine match_batch__fixed(pat, strs, res, status)
      type(fx_pattern_t), intent(in) :: pat
      character(len=*), intent(in), contiguous :: strs(:)
      logical, intent(out) :: res(:)
      integer, intent(out) :: status
      integer(c_int8_t), allocatable :: out(:)
      allocate(out(size(strs)))
      status = fx_match_fixed(pat%handle, strs, int(size(strs), c_int64_t), int(len(strs), c_int64_t), out)
      res = out /= 0
   end subroutine match_batch__fixed

   subroutine in_batch__fixed(pat, strs, res, status)
      type(fx_pattern_t), intent(in) :: pat
      character(len=*), intent(in), contiguous :: strs(:)
      logical, intent(out) :: res(:)
      integer, intent(out) :: status
      integer(c_int8_t), allocatable :: out(:)
      allocate(out(size(strs)))
      status = fx_in_fixed(pat%handle, strs, int(size(strs), c_int64_t), int(len(strs), c_int64_t), out)
      res = out /= 0
   end subroutine in_batch__fixed

   ! variable-length strings: flat buffer + offsets(0:n), 0-based byte offsets, offsets(0) = 0
   subroutine match_batch__ragged(pat, buf, offsets, res, status)
      type(fx_pattern_t), intent(in) :: pat
      character(len=*), intent(in) :: buf
      integer(int64), intent(in) :: offsets(0:)
      logical, intent(out) :: res(:)
      integer, intent(out) :: status
      integer(c_int8_t), allocatable :: out(:)
      allocate(out(size(res)))
      status = fx_match_batch(pat%handle, buf, offsets, int(size(res), c_int64_t), out)
      res = out /= 0
   end subroutine match_batch__ragged

   subroutine in_batch__ragged(pat, buf, offsets, res, status)
      type(fx_pattern_t), intent(in) :: pat
      character(len=*), intent(in) :: buf
      integer(int64), intent(in) :: offsets(0:)
      logical, intent(out) :: res(:)
      integer, intent(out) :: status
      integer(c_int8_t), allocatable :: out(:)
      allocate(out(size(res)))
      status = fx_in_batch(pat%handle, buf, offsets, int(size(res), c_int64_t), out)
      res = out /= 0
   end subroutine in_batch__ragged

   subroutine regex_batch(pat, buf, offsets, from, to, status)
      type(fx_pattern_t), intent(in) :: pat
      character(len=*), intent(in) :: buf
      integer(int64), intent(in) :: offsets(0:)
      integer(int64), intent(out) :: from(:), to(:)    ! 1-based inclusive, relative to each string; 0/0 = no match
      integer, intent(out) :: status
      status = fx_regex_batch(pat%handle, buf, offsets, int(size(from), c_int64_t), from, to)
   end subroutine regex_batch

   subroutine regex_buffer(pat, buf, from, to, status)
      type(fx_pattern_t), intent(in) :: pat
      character(len=*), intent(in) :: buf
      integer(int64), intent(out) :: from, to           ! 64-bit: the reference's default-integer indices stop at 2 GiB
      integer, intent(out) :: status
      status = fx_regex_buffer(pat%handle, buf, int(len(buf, kind=int64), c_int64_t), from, to)
   end subroutine regex_buffer

end module forgex_b200_m

! forgex_b200_tables_m -- the pattern compile on the FORTRAN side (SURVEY.md 8f-1, hard part H1).
!
! Instead of re-implementing Forgex's front end, this module REUSES it: tree%build parses the pattern, extract_literal
! produces the three literals, automaton%preprocess / %init build the NFA, and the lazy DFA is then explored EAGERLY --
! a breadth-first search that calls the reference's own automaton%construct (src/automaton_m.F90:333) for every DFA
! state and one representative symbol of every segment of the disjoint alphabet (automaton%all_segments,
! src/automaton_m.F90:34).  What comes out is the anchored code-point DFA; fx_compile_from_dfa (C ABI) derives every
! device table from it.  Language and quirks are therefore Forgex's own by construction; the C++ half of this route is
! tested without a Fortran compiler (tests/test_host_tables.py::test_fortran_side_route_from_an_anchored_dfa feeds it
! the DFA the product's own front end produces and checks every entry point against the oracle).
!
! STATUS: UNVERIFIED -- never compiled (no Fortran compiler in this image).  CI note for a maintainer: build with
!   cmake -DFORGEX_B200_FORTRAN=ON (CMakeLists.txt) next to the reference's sources and run fortran/smoke.F90, which
!   compares compile_with_forgex_front_end + match_batch against operator(.match.) on the reference's own test vectors.
module forgex_b200_tables_m
   use, intrinsic :: iso_c_binding
   use, intrinsic :: iso_fortran_env, only: int32
   use :: forgex_automaton_m, only: automaton_t
   use :: forgex_syntax_tree_graph_m, only: tree_t
   use :: forgex_syntax_tree_optimize_m, only: extract_literal
   use :: forgex_segment_m, only: segment_t
   use :: forgex_utf8_m, only: char_utf8
   use :: forgex_parameters_m, only: DFA_INVALID_INDEX, DFA_STATE_HARD_LIMIT, UTF8_CODE_MAX
   use :: forgex_b200_m, only: fx_pattern_t, fx_compile_from_dfa, FX_ERR_DFA_STATE_CAP
   implicit none
   private
   public :: compile_with_forgex_front_end

contains

   !> `pattern` must already carry the entry point's own preprocessing: trim(pattern) for `.in.` / regex
   !> (src/forgex.F90:95, :260), the caret / dollar stripping of operator__match (src/forgex.F90:182-190).
   subroutine compile_with_forgex_front_end(pattern, op, pat, status, state_cap)
      character(*), intent(in) :: pattern
      integer(c_int), intent(in) :: op
      type(fx_pattern_t), intent(inout) :: pat
      integer, intent(out) :: status
      integer, intent(in), optional :: state_cap

      type(tree_t) :: tree
      type(automaton_t) :: automaton
      character(:), allocatable :: all, prefix, suffix, factor, buff
      integer(c_int32_t), allocatable :: cuts(:), delta(:, :)
      integer(c_int8_t), allocatable :: accept(:)
      integer(int32), allocatable :: seg_class(:)
      integer(int32) :: cap, nseg, ncls, i, s, c, dst, top, q0, lo, hi, prev_hi

      cap = DFA_STATE_HARD_LIMIT - 1
      if (present(state_cap)) cap = state_cap
      all = ''; prefix = ''; suffix = ''; factor = ''
      buff = pattern
      call tree%build(buff)
      if (.not. tree%is_valid) then
         status = tree%code                    ! SYNTAX_* code, as `regex` reports it (src/forgex.F90:266-274)
         return
      end if
      call extract_literal(tree, all, prefix, suffix, factor)
      call automaton%preprocess(tree)
      call automaton%init()
      q0 = automaton%initial_index

      ! ---- the alphabet: Forgex's disjoint segments in ascending order, the gaps between them are dead classes ----
      nseg = 0
      do i = 1, size(automaton%all_segments)
         if (automaton%all_segments(i)%min <= UTF8_CODE_MAX .and. automaton%all_segments(i)%min >= 0) nseg = nseg + 1
      end do
      allocate(cuts(0:2*nseg + 1), seg_class(size(automaton%all_segments)))
      ncls = 0
      prev_hi = -1
      cuts(0) = 0
      seg_class = -1
      do i = 1, size(automaton%all_segments)     ! (disjoin_nfa leaves them sorted; a maintainer should assert it)
         lo = automaton%all_segments(i)%min
         hi = automaton%all_segments(i)%max
         if (lo > UTF8_CODE_MAX .or. lo < 0) cycle
         if (lo > prev_hi + 1) then              ! a gap in front of this segment: one dead class
            ncls = ncls + 1
            cuts(ncls) = lo
         end if
         seg_class(i) = ncls                     ! 0-based class of this segment: [cuts(ncls), hi]
         ncls = ncls + 1
         cuts(ncls) = hi + 1
         prev_hi = hi
      end do
      ! (cuts(0:ncls) now bounds ncls classes; code points above the last cut match nothing)

      ! ---- breadth-first search over the reference's own per-symbol step ----
      allocate(delta(0:ncls - 1, 0:cap), accept(0:cap))
      delta = 0
      accept = 0
      s = q0
      do while (s <= automaton%dfa%dfa_top)
         if (automaton%dfa%dfa_top > cap) then
            status = FX_ERR_DFA_STATE_CAP
            call automaton%free()
            return
         end if
         do i = 1, size(automaton%all_segments)
            c = seg_class(i)
            if (c < 0) cycle
            call automaton%construct(s, dst, char_utf8(automaton%all_segments(i)%min))   ! src/automaton_m.F90:333
            if (dst /= DFA_INVALID_INDEX) delta(c, s) = dst
         end do
         s = s + 1
      end do
      top = automaton%dfa%dfa_top
      do s = q0, top
         if (automaton%dfa%nodes(s)%accepted) accept(s) = 1
      end do

      ! rows 0 .. top, state 0 = dead (DFA_INVALID_INDEX = 0, src/essential/parameters_m.f90:133); row-major for C
      status = fx_compile_from_dfa(op, cuts, ncls, reshape(delta(:, 0:top), [ncls * (top + 1)]), top + 1, accept, q0, &
                                   all, int(len(all), c_int64_t), prefix, int(len(prefix), c_int64_t), &
                                   suffix, int(len(suffix), c_int64_t), pat%handle)
      pat%status = status
      call automaton%free()
   end subroutine compile_with_forgex_front_end

end module forgex_b200_tables_m

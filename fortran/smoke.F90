! smoke.F90 -- what CI would run once a Fortran compiler exists (cmake -DFORGEX_B200_FORTRAN=ON): the batch forms and
! the Fortran-side compile route against the reference's own operators, on vectors of test/test_api/test_case_003.f90.
! STATUS: UNVERIFIED -- never compiled (no Fortran compiler in this image).
program forgex_b200_smoke
   use, intrinsic :: iso_c_binding
   use :: forgex, only: operator(.match.), operator(.in.)
   use :: forgex_b200_m
   use :: forgex_b200_tables_m, only: compile_with_forgex_front_end
   implicit none
   character(len=8) :: strs(4)
   logical :: hit(4), ref(4)
   type(fx_pattern_t) :: a, b
   integer :: status, i

   strs = [character(len=8) :: '123-4567', '12a-4567', '000-0000', '123-456']
   do i = 1, 4
      ref(i) = '\d{3}-\d{4}' .match. strs(i)            ! the reference, element by element
   end do
   call a%compile('\d{3}-\d{4}', FX_OP_MATCH)             ! host-side compile in the library (C++)
   call match_batch(a, strs, hit, status)
   if (status /= 0 .or. any(hit .neqv. ref)) error stop 'match_batch differs from operator(.match.)'
   call compile_with_forgex_front_end('\d{3}-\d{4}', FX_OP_MATCH, b, status)    ! Forgex's own front end + eager BFS
   if (status /= 0) error stop 'compile_with_forgex_front_end failed'
   call match_batch(b, strs, hit, status)
   if (status /= 0 .or. any(hit .neqv. ref)) error stop 'Fortran-side tables differ from operator(.match.)'
   if (fx_in_value('foo(bar|baz)', 12_c_int64_t, 'xxfoobaz', 8_c_int64_t) /= 1) error stop 'fx_in_value'
   if (('foo(bar|baz)' .in. 'xxfoobaz') .neqv. .true.) error stop 'reference .in.'
   call a%free()
   call b%free()
   print *, 'forgex_b200 fortran smoke: ok'
end program forgex_b200_smoke

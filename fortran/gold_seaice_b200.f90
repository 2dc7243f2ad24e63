! gold_seaice_b200.f90 -- MODULE gold_seaice_b200: step_seaice with the reference's name and argument list
! (src/goldsteinseaice/gold_seaice.f90:511-520).  MODULE gold_seaice stays in the build; fortran/use_b200.py switches
! gold_seaice_wrapper's USE line (genie_loop_wrappers.f90:94-95).
MODULE gold_seaice_b200
  USE, INTRINSIC :: ISO_C_BINDING
  USE cgenie_b200_c
  USE gold_seaice_lib, ONLY: maxi, maxj, iwstp, itstp, npstp
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: step_seaice

CONTAINS

  SUBROUTINE step_seaice(istep, dhght_sic, dfrac_sic, ustar_ocn, vstar_ocn, &
       & hght_sic, frac_sic, temp_sic, albd_sic, sic_FW_ocn, sic_FX0_ocn, &
       & test_energy_seaice, test_water_seaice)
    REAL :: dhght_sic(maxi,maxj), dfrac_sic(maxi,maxj), ustar_ocn(maxi,maxj), vstar_ocn(maxi,maxj), &
         & temp_sic(maxi,maxj), albd_sic(maxi,maxj)
    REAL, TARGET :: hght_sic(maxi,maxj), frac_sic(maxi,maxj), sic_FW_ocn(maxi,maxj), sic_FX0_ocn(maxi,maxj)
    INTEGER :: istep
    REAL :: test_energy_seaice, test_water_seaice
    TYPE(cg_seaice_io), TARGET :: io
    INTEGER(C_INT) :: rc
    CALL cg_ensure_handle()
    IF (MOD(istep, itstp) == 0 .OR. MOD(istep, iwstp) == 0 .OR. MOD(istep, npstp) == 0) THEN
       io%hght_sic = C_LOC(hght_sic) ; io%frac_sic = C_LOC(frac_sic)
       io%waterflux_ocn = C_LOC(sic_FW_ocn) ; io%conductflux_ocn = C_LOC(sic_FX0_ocn)
       rc = cg_seaice_step(cg_h, INT(istep, C_INT), C_LOC(io))
    ELSE
       rc = cg_seaice_step(cg_h, INT(istep, C_INT), C_NULL_PTR)
    END IF
    CALL cg_check(rc, 'cg_seaice_step')
  END SUBROUTINE step_seaice

END MODULE gold_seaice_b200
! Stand-ins for mpi_utils (one MPI rank: a reduction is a copy) and model_utils (handle_err only).
module mpi_utils
  use nrtype
  implicit none
  public :: shr_mpi_reduce
contains
  subroutine shr_mpi_reduce(sendArray, op, recvArray, ierr, message)
    real(dp),     intent(in)  :: sendArray(:)
    character(*), intent(in)  :: op
    real(dp),     intent(out) :: recvArray(:)
    integer(i4b), intent(out) :: ierr
    character(*), intent(out) :: message
    ierr = 0; message = 'shr_mpi_reduce/'//trim(op)
    recvArray = sendArray
  end subroutine shr_mpi_reduce
end module mpi_utils

module model_utils
  use nrtype
  implicit none
  public :: handle_err
contains
  subroutine handle_err(err, message)
    integer(i4b), intent(in) :: err
    character(*), intent(in) :: message
    if (err /= 0) then
      write(*, '(A)') 'FATAL ERROR: '//trim(message)
      stop 1
    end if
  end subroutine handle_err
end module model_utils

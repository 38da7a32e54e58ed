! Stand-in for the PIO library module: globalData only declares variables of these two types.
module pio
  implicit none
  type, public :: iosystem_desc_t
    integer :: dummy = 0
  end type iosystem_desc_t
  type, public :: io_desc_t
    integer :: dummy = 0
  end type io_desc_t
  type, public :: file_desc_t
    integer :: dummy = 0
  end type file_desc_t
end module pio

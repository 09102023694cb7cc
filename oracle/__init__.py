"""TEST INFRASTRUCTURE ONLY: CPU oracle for the Struct3d native path.

`oracle.cport`   -- ctypes binding of the plain-C restatement (oracle/s3d_oracle.c).
`oracle.refshim` -- ctypes binding of the reference's own sources compiled unchanged
                    (oracle/_ref/libs3dref.so; exists only where /root/reference was present
                    at build time, or where the prebuilt .so travelled with the repo).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package struct3d_b200 never does.
"""

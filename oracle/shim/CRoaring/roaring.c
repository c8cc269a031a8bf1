/* TEST INFRASTRUCTURE ONLY — empty stand-in for the amalgamated CRoaring C source that
 * /root/reference/src/Parser.cpp:8 textually includes.  The set type itself is header-only,
 * see roaring.hh next to this file. */

"""TEST INFRASTRUCTURE — CPU oracle (see oracle/oracle.hpp). Not part of the product path."""

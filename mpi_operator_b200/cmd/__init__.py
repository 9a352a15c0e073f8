"""Process shells: the operator daemon (main/options/server) and mpijobctl."""

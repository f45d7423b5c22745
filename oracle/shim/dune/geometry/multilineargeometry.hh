// TEST INFRASTRUCTURE ONLY (oracle/): empty stand-in (header is included but nothing from it is instantiated).
#pragma once

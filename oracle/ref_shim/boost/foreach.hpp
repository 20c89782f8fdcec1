#pragma once
#define BOOST_FOREACH(decl, range) for (decl : range)

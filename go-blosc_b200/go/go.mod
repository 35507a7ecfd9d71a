module github.com/example/go-blosc-b200

go 1.22
